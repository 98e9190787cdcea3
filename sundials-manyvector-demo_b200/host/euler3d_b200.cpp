// ---------------------------------------------------------------------------
// euler3d_b200.cpp -- native explicit driver for the B200 fluid RHS (SURVEY.md 8(f-1), 8(f-2),
// 8(f-4)).  Plays the part of src/euler3D_main.cpp for the explicit problems of the
// reference when SUNDIALS is not available: same input files ("key = value" lines and
// --key=value overrides, io.cpp:225-367), same run structure (initial outputs, nout
// evolve/diagnostics cycles, final statistics and conservation check,
// euler3D_main.cpp:300-462), same diagnostics text (errI / errR, stats table, conservation).
//
//   euler3d_b200 --problem=sod_x -f input_sod.txt [--nx=400 --rtol=1e-6 ...]
//
// Problems (chosen at link time in the reference, by name here; host/problems.hpp): sod_{x,y,z},
// linear_advection_{x,y,z}, rayleigh_taylor, hurricane_{xy,yz,zx}, fluid_blast, primordial_blast
// (fluid + passive tracers).  --nchem=<int> plays the part of the reference's compile-time NVAR
// (nchem = NVAR - 5, euler3D.hpp:52-59).  --output=1 writes output-<iout>.eb200 files (the flat
// stand-in for output_solution's HDF5 files), --restart=<iout> starts from one (io.cpp:940).
// The state lives on the GPU for the whole run; this file contains no CUDA: device memory,
// the right-hand side, the stage combinations and the error norm all go through the C ABI
// of include/eulerb200.h.  The time integrator is the embedded explicit Runge-Kutta loop
// described in driver.py (ARKODE's default tables and controller constants; it is not
// ARKODE).  One rank.
// ---------------------------------------------------------------------------
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>
#include "eulerb200.h"
#include "problems.hpp"
#include "erk_tables.hpp"
#include "erk_stepper.hpp"

namespace {

using eb_problems::Problem;

// Cumulative wall-clock timers of the run regions the reference profiles (class Profile,
// profiler.hpp:50-108; slots of euler3D.hpp:97-118 that exist here).  The RHS and stability calls of
// this driver return after the device has finished, so host timers around them are meaningful.
struct Timer {
  double total = 0, t0 = -1;
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  void start() { t0 = now(); }
  void stop() { if (t0 >= 0) total += now() - t0; t0 = -1; }
  void print(const char* name) const { printf("Total %s time = \t%.2e  ( min / max  =  %.2e / %.2e )\n", name, total, total, total); }
};
Timer g_prof_rhs, g_prof_stab;

struct Inputs {
  std::map<std::string, double> v;
  std::string problem = "sod_x";
  double get(const std::string& k, double dflt) const { auto it = v.find(k); return it == v.end() ? dflt : it->second; }
};

bool parse_line(const std::string& line, Inputs& in)
{
  std::string s = line.substr(0, line.find('#'));
  const size_t eq = s.find('=');
  if (eq == std::string::npos) return false;
  auto trim = [](std::string t) {
    const size_t a = t.find_first_not_of(" \t\r\n"), b = t.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
  };
  const std::string key = trim(s.substr(0, eq)), val = trim(s.substr(eq + 1));
  if (key.empty() || val.empty()) return false;
  if (key == "problem") { in.problem = val; return true; }
  char* end = nullptr;
  const double x = strtod(val.c_str(), &end);
  if (end == val.c_str()) return false;
  in.v[key] = x;
  return true;
}

struct Vec {                       // the MPIManyVector composition: 5 fluid sub-vectors (+ chem)
  double* sub[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  long len[6] = {0, 0, 0, 0, 0, 0};
  int nsub = 5;
};

Vec new_vec(long N, int nchem)
{
  Vec v;
  v.nsub = 5 + (nchem > 0 ? 1 : 0);
  for (int f = 0; f < v.nsub; f++) {
    v.len[f] = f < 5 ? N : N * nchem;
    v.sub[f] = (double*)eulerb200_device_alloc(sizeof(double) * v.len[f]);
    if (!v.sub[f]) { fprintf(stderr, "MEMORY_ERROR: device allocation failed\n"); exit(1); }
  }
  return v;
}
void free_vec(Vec& v) { for (int f = 0; f < v.nsub; f++) eulerb200_device_free(v.sub[f]); }

// Vector operations of the shared ERK loop (host/erk_stepper.hpp) on device-resident vectors:
// everything goes through the C ABI.
struct DeviceOps {
  typedef ::Vec Vec;
  eulerb200_ctx* ctx = NULL;
  long nglobal = 0;
  void die(const char* what) { fprintf(stderr, "\n%s: %s\n\n", what, eulerb200_last_error(ctx)); exit(1); }
  void lincomb(Vec& out, int n, const double* c, Vec* const* v)
  {
    for (int f = 0; f < out.nsub; f++) {
      const double* x[16];
      for (int q = 0; q < n; q++) x[q] = v[q]->sub[f];
      if (eulerb200_vec_lincomb(ctx, n, c, x, out.sub[f], out.len[f], NULL)) die("eulerb200_vec_lincomb");
    }
  }
  double wrms(const Vec& x, const Vec& y, double rtol, double atol)
  {
    double r = 0;
    if (eulerb200_vec_wrms(ctx, x.sub, y.sub, rtol, atol, nglobal, &r, NULL)) die("eulerb200_vec_wrms");
    return r;
  }
  int rhs(double tt, Vec& y, Vec& out)
  {
    g_prof_rhs.start();
    if (eulerb200_rhs(ctx, tt, y.sub, out.sub, NULL)) die("fEuler");
    g_prof_rhs.stop();
    return 0;
  }
  int stability(Vec& w, double, double cfl, double* dt)
  {
    g_prof_stab.start();
    if (eulerb200_stability(ctx, w.sub, cfl, dt, NULL)) die("stability");
    g_prof_stab.stop();
    return 0;
  }
};
typedef ErkStepper<DeviceOps> Stepper;

}  // namespace

// --help: what this driver understands of the reference's input keys (io.cpp:120-222 prints the
// reference's own list; keys of the implicit / multirate solvers have no meaning here)
static void usage()
{
  printf("euler3d_b200 -f <input file> [--key=value ...]      (keys as in the reference's input files)\n\n"
         "problem     --problem=<name>   sod_x|sod_y|sod_z, linear_advection_x|_y|_z, rayleigh_taylor,\n"
         "                               hurricane_xy|_yz|_zx, fluid_blast, primordial_blast\n"
         "            --nchem=<int>      advected species (the reference's compile-time NVAR - 5), 0..64\n"
         "grid        --nx --ny --nz, --xl --xr --yl --yr --zl --zr, --gamma\n"
         "            --xlbc --xrbc --ylbc --yrbc --zlbc --zrbc   0 periodic, 1 Neumann, 2 Dirichlet, 3 reflecting\n"
         "units       --MassUnits --LengthUnits --TimeUnits\n"
         "run         --t0 --tf --nout, --showstats=1, --output=1 (write output-<n>.eb200), --restart=<n>\n"
         "stepping    --order=2|3|4|5|6|8  or  --order=0 --etable=0|1|3|6|7|8|10|11|12   (order overrides etable)\n"
         "            --rtol --atol --fixedstep=1 --hmax (the fixed step) --hmin --h0 --cfl --mxsteps --maxnef\n"
         "            --fixedstep=1 --htrans=<t>  adaptive (steps <= hmax) over (t0, t0+htrans], then fixed steps of hmax\n"
         "            --safety --bias --growth --k1 --k2 --k3 --etamx1 --etamxf   (0: ARKODE's default)\n");
}

int main(int argc, char** argv)
{
  Timer prof_setup, prof_io, prof_trans, prof_sim;
  prof_setup.start();
  Inputs in;
  std::vector<std::string> overrides;
  for (int a = 1; a < argc; a++) {
    const std::string s = argv[a];
    if (s == "--help" || s == "-h") { usage(); return 0; }
    if (s == "-f" && a + 1 < argc) {
      std::ifstream fin(argv[++a]);
      if (!fin) { fprintf(stderr, "cannot open input file %s\n", argv[a]); return 1; }
      std::string line;
      while (std::getline(fin, line)) parse_line(line, in);
    } else if (s.compare(0, 2, "--") == 0) overrides.push_back(s.substr(2));
  }
  for (const auto& o : overrides) parse_line(o, in);          // command line wins (io.cpp:310-367)

  Problem P;
  P.name = in.problem;
  P.nx = (long)in.get("nx", 3); P.ny = (long)in.get("ny", 3); P.nz = (long)in.get("nz", 3);
  P.xl = in.get("xl", 0); P.xr = in.get("xr", 1); P.yl = in.get("yl", 0); P.yr = in.get("yr", 1);
  P.zl = in.get("zl", 0); P.zr = in.get("zr", 1);
  P.gamma = in.get("gamma", 1.4);
  P.nchem = (int)in.get("nchem", 0);
  P.MassUnits = in.get("MassUnits", 1.0); P.LengthUnits = in.get("LengthUnits", 1.0); P.TimeUnits = in.get("TimeUnits", 1.0);
  double t0 = in.get("t0", 0.0);
  const double tf = in.get("tf", 1.0);
  const int nout = (int)in.get("nout", 10), showstats = (int)in.get("showstats", 0);
  const int write_files = (int)in.get("output", 0), restart = (int)in.get("restart", -1);
  if (P.nchem < 0 || P.nchem > 64) { fprintf(stderr, "illegal nchem = %d\n", P.nchem); return 1; }
  Table table;
  if (!make_table((int)in.get("order", 4), (int)in.get("etable", -1), table)) {
    fprintf(stderr, "\nERROR: no explicit Butcher table for order = %d / etable = %d (orders 2-6, 8; table ids 0 1 3 6 7 8 10 11 12)\n\n",
            (int)in.get("order", 4), (int)in.get("etable", -1));
    return 1;
  }
  if (!table.embedded && ((int)in.get("fixedstep", 0) == 0 || in.get("htrans", 0) > 0)) {
    fprintf(stderr, "\nERROR: this Butcher table has no embedding: it needs fixedstep = 1\n\n");
    return 1;
  }

  eulerb200_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.nxl = P.nx; cfg.nyl = P.ny; cfg.nzl = P.nz;
  cfg.nchem = P.nchem; cfg.device = -1;
  cfg.dx = P.dx(); cfg.dy = P.dy(); cfg.dz = P.dz();
  cfg.gamma = P.gamma;
  const char* bcn[6] = {"xlbc", "xrbc", "ylbc", "yrbc", "zlbc", "zrbc"};
  for (int f = 0; f < 6; f++) {
    cfg.bc[f] = (int)in.get(bcn[f], 0);
    cfg.nbr[f] = cfg.bc[f] == EULERB200_BC_PERIODIC ? 0 : EULERB200_NO_NEIGHBOR;
  }
  cfg.rank = 0; cfg.nranks = 1;
  if (P.name == "rayleigh_taylor") cfg.forcing[2] = -0.1;     // rayleigh_taylor.cpp:117-128
  eulerb200_ctx* ctx = NULL;
  if (eulerb200_create(&cfg, &ctx) != 0) {
    fprintf(stderr, "\neulerb200_create failed: %s\n\n", eulerb200_last_error(NULL));
    return 1;
  }

  printf("\n3D compressible inviscid Euler test problem (B200 fluid RHS): %s\n", P.name.c_str());
  printf("   spatial domain: [%g, %g] x [%g, %g] x [%g, %g]\n", P.xl, P.xr, P.yl, P.yr, P.zl, P.zr);
  printf("   time domain = (%g, %g]\n", t0, tf);
  printf("   bdry cond (0=per, 1=Neu, 2=Dir, 3=refl): [%d, %d] x [%d, %d] x [%d, %d]\n",
         cfg.bc[0], cfg.bc[1], cfg.bc[2], cfg.bc[3], cfg.bc[4], cfg.bc[5]);
  printf("   gamma: %g\n   spatial grid: %ld x %ld x %ld\n", P.gamma, P.nx, P.ny, P.nz);
  if (P.nchem > 0) printf("   num chemical species: %d\n", P.nchem);

  const long N = P.nx * P.ny * P.nz;
  Stepper S;
  S.ops.ctx = ctx;
  S.T = table;
  S.ops.nglobal = (5 + P.nchem) * N;
  S.rtol = in.get("rtol", 1e-8); S.atol = in.get("atol", 1e-12);
  // fixedstep = 0 adaptive, 1 fixed steps of hmax; with htrans > 0 as well: adaptive (steps <= hmax)
  // over (t0, t0+htrans], fixed afterwards (euler3D_main.cpp:79-88,221,345-367)
  const double htrans = in.get("htrans", 0);
  int fixed_mode = (int)in.get("fixedstep", 0);
  if (fixed_mode && in.get("hmax", 0) <= 0) {
    fprintf(stderr, "\nError: fixed time stepping requires hmax > 0 (%g given)\n", in.get("hmax", 0));
    return 1;
  }
  if (fixed_mode && htrans > 0) fixed_mode = 2;
  S.fixedstep = fixed_mode == 1 ? 1 : 0;
  S.hmin = in.get("hmin", 0); S.hmax = in.get("hmax", 0); S.h0 = in.get("h0", 0);
  S.cfl = in.get("cfl", 0);
  S.mxsteps = (int)in.get("mxsteps", 5000);
  if (S.mxsteps <= 0) S.mxsteps = 500;                        // 0 => ARKODE's default (MXSTEP_DEFAULT)
  auto dflt = [&](const char* k, double d) { const double x = in.get(k, 0); return x != 0 ? x : d; };
  S.safety = dflt("safety", 0.96); S.bias = dflt("bias", 1.5); S.growth = dflt("growth", 20.0);
  S.k1 = dflt("k1", 0.58); S.k2 = dflt("k2", 0.21); S.k3 = dflt("k3", 0.1);
  S.etamx1 = dflt("etamx1", 1e4); S.etamxf = dflt("etamxf", 0.3);
  S.maxnef = (int)dflt("maxnef", 7);
  S.t = t0;
  S.w = new_vec(N, P.nchem); S.ytmp = new_vec(N, P.nchem); S.yerr = new_vec(N, P.nchem);
  for (int i = 0; i < S.T.s; i++) S.k[i] = new_vec(N, P.nchem);

  // initial conditions or restart file (host, then one copy to the device)
  std::vector<std::vector<double>> host(5, std::vector<double>(N));
  std::vector<double> host_chem((size_t)N * P.nchem);
  double* const hf[5] = {host[0].data(), host[1].data(), host[2].data(), host[3].data(), host[4].data()};
  bool analytic = true;
  if (eb_problems::initial_conditions(P, t0, hf, host_chem.data(), &analytic) != 0) return 1;
  if (restart >= 0) {
    if (eb_problems::read_solution(eb_problems::solution_name(restart), P, &t0, hf, host_chem.data()) != 0) return 1;
    printf("   restarting from %s at t = %g\n", eb_problems::solution_name(restart).c_str(), t0);
    S.t = t0;
  }
  for (int f = 0; f < 5; f++) eulerb200_copy_to_device(S.w.sub[f], host[f].data(), sizeof(double) * N);
  if (P.nchem > 0) eulerb200_copy_to_device(S.w.sub[5], host_chem.data(), sizeof(double) * N * P.nchem);

  double mass0 = -1, energy0 = -1;
  int iout_file = restart >= 0 ? restart : 0, outputs_done = 0;
  // the three per-output actions, called in the reference's order (euler3D_main.cpp:304-332,405-424)
  auto fetch = [&]() { for (int f = 0; f < 5; f++) eulerb200_copy_to_host(host[f].data(), S.w.sub[f], sizeof(double) * N); };
  auto write_file = [&](double t) {
    if (write_files) {                                        // output_solution, io.cpp:716-930
      if (P.nchem > 0) eulerb200_copy_to_host(host_chem.data(), S.w.sub[5], sizeof(double) * N * P.nchem);
      if (eb_problems::write_solution(eb_problems::solution_name(iout_file), P, t, hf, host_chem.data()) != 0) {
        fprintf(stderr, "output_solution: cannot write %s\n", eb_problems::solution_name(iout_file).c_str());
        exit(1);
      }
      // write_parameters (io.cpp:645-714): an input file that continues this run from the file just
      // written -- "euler3d_b200 -f restart_parameters.txt"
      std::ofstream pf("restart_parameters.txt");
      pf << "# euler3d_b200 restart file\nproblem = " << P.name << "\n";
      std::map<std::string, double> vals = in.v;
      vals["t0"] = t; vals["nout"] = nout - outputs_done; vals["h0"] = S.h; vals["restart"] = iout_file; vals["output"] = 1;
      char buf[64];
      for (const auto& kv : vals) { snprintf(buf, sizeof buf, "%.17g", kv.second); pf << kv.first << " = " << buf << "\n"; }
      iout_file++;
    }
  };
  auto diagnostics = [&](double t) {
    if (P.name.compare(0, 9, "hurricane") == 0) {             // hurricane.cpp:217-334: rho and the momenta
      double errI[4] = {0, 0, 0, 0}, errR[4] = {0, 0, 0, 0};
      for (long k = 0; k < P.nz; k++)
        for (long j = 0; j < P.ny; j++)
          for (long i = 0; i < P.nx; i++) {
            double w4[4];
            eb_problems::hurricane_true(P, t, i, j, k, w4);
            const long c = i + P.nx * (j + P.ny * k);
            for (int f = 0; f < 4; f++) {
              const double e = fabs(w4[f] - host[f][c]);
              errI[f] = std::max(errI[f], e); errR[f] += e * e;
            }
          }
      printf("     errI = %9.2e  %9.2e  %9.2e  %9.2e\n", errI[0], errI[1], errI[2], errI[3]);
      printf("     errR = %9.2e  %9.2e  %9.2e  %9.2e\n", sqrt(errR[0] / N), sqrt(errR[1] / N), sqrt(errR[2] / N), sqrt(errR[3] / N));
    } else if (analytic) {                                           // output_diagnostics of the problem file
      double errI[5] = {0, 0, 0, 0, 0}, errR[5] = {0, 0, 0, 0, 0};
      for (long k = 0; k < P.nz; k++)
        for (long j = 0; j < P.ny; j++)
          for (long i = 0; i < P.nx; i++) {
            double w5[5];
            eb_problems::state_at(P, t, i, j, k, w5);
            const long c = i + P.nx * (j + P.ny * k);
            for (int f = 0; f < 5; f++) {
              const double e = fabs(w5[f] - host[f][c]);
              errI[f] = std::max(errI[f], e); errR[f] += e * e;
            }
          }
      printf("     errI = %9.2e  %9.2e  %9.2e  %9.2e  %9.2e\n", errI[0], errI[1], errI[2], errI[3], errI[4]);
      printf("     errR = %9.2e  %9.2e  %9.2e  %9.2e  %9.2e\n", sqrt(errR[0] / N), sqrt(errR[1] / N),
             sqrt(errR[2] / N), sqrt(errR[3] / N), sqrt(errR[4] / N));
    }
  };
  auto stats = [&](double t, int firstlast) {
    if (showstats) {                                          // print_stats (CGS values), io.cpp:552-636
      const double su[5] = {P.DensityUnits(), P.MomentumUnits(), P.MomentumUnits(), P.MomentumUnits(), P.EnergyUnits()};
      if (firstlast == 0) {
        printf("\n      t       ||rho||   ||mx||    ||my||    ||mz||    ||et||   ");
        for (int v = 0; v < P.nchem; v++) printf(" ||c%d||   ", v);
        printf("   nst\n");
      }
      if (firstlast != 1) {
        printf("   ------------------------------------------------------------");
        for (int v = 0; v < P.nchem; v++) printf("----------");
        printf("-------\n");
      }
      if (firstlast == 2) return;
      printf("  %9.1e", t);
      for (int f = 0; f < 5; f++) {
        double s = 0;
        for (long c = 0; c < N; c++) s += (host[f][c] * su[f]) * (host[f][c] * su[f]);
        printf(" %9.1e", sqrt(s / N));
      }
      if (P.nchem > 0) {
        // (always: write_file() runs AFTER stats() in the output loop, its copy is one output old here)
        eulerb200_copy_to_host(host_chem.data(), S.w.sub[5], sizeof(double) * N * P.nchem);
        for (int v = 0; v < P.nchem; v++) {
          double s = 0;
          for (long c = 0; c < N; c++) s += host_chem[c * P.nchem + v] * host_chem[c * P.nchem + v];
          printf(" %9.1e", sqrt(s / N));
        }
      }
      printf("  %6ld\n", S.nst);
    }
  };
  auto conservation = [&]() {                                 // check_conservation, io.cpp:504-541
    for (int f = 0; f < 5; f += 4) eulerb200_copy_to_host(host[f].data(), S.w.sub[f], sizeof(double) * N);
    double m = 0, e = 0;
    for (long c = 0; c < N; c++) { m += host[0][c]; e += host[4][c]; }
    const double vol = P.dx() * P.dy() * P.dz() * pow(P.LengthUnits, 3);      // CGS totals, io.cpp:522-524
    m *= vol * P.DensityUnits(); e *= vol * P.EnergyUnits();
    if (mass0 == -1) { printf("   Total mass   = %.16e\n   Total energy = %.16e\n", m, e); mass0 = m; energy0 = e; }
    else {
      printf("   Mass conservation relative change   = %7.2e\n", fabs(m - mass0) / mass0);
      printf("   Energy conservation relative change = %7.2e\n", fabs(e - energy0) / energy0);
    }
  };

  prof_io.start();
  printf("\nWriting initial batch of outputs\n");
  if (showstats) conservation();
  fetch();
  write_file(t0);
  stats(t0, 0);
  diagnostics(t0);
  prof_io.stop();
  prof_setup.stop();

  if (fixed_mode == 2) {               // initial transient (euler3D_main.cpp:340-367)
    prof_trans.start();
    if (S.evolve(t0 + htrans) != 0) { fprintf(stderr, "Solver failure, stopping integration\n"); return 1; }
    S.fixedstep = 1; S.h = 0.0;         // ARKStepSetFixedStep(hmax)
    prof_trans.stop();
  }
  prof_sim.start();
  const double dTout = (tf - t0) / nout;
  double tout = t0 + dTout;
  for (int iout = 0; iout < nout; iout++) {
    if (S.evolve(tout) != 0) { fprintf(stderr, "Solver failure, stopping integration\n"); return 1; }
    outputs_done++;
    prof_io.start();
    fetch();
    diagnostics(S.t);
    stats(S.t, 1);
    write_file(S.t);
    prof_io.stop();
    tout = std::min(tout + dTout, tf);
  }
  stats(S.t, 2);
  prof_sim.stop();

  printf("\nFinal Solver Statistics:\n");
  printf("   Internal solver steps = %ld (attempted = %ld)\n", S.nst, S.nst_a);
  printf("   Total RHS evals:  Fe = %ld,  Fi = 0\n", S.nfe);
  printf("   Total number of error test failures = %ld\n", S.netf);
  printf("   GPU kernel launches = %lld\n", (long long)eulerb200_launch_count(ctx));
  printf("\nProfiling Results:\n");                         // euler3D_main.cpp:449-461 (the slots that exist here)
  prof_setup.print("setup"); prof_io.print("I/O"); g_prof_rhs.print("RHS"); g_prof_stab.print("dt_stab");
  prof_trans.print("trans"); prof_sim.print("sim");
  if (showstats) { printf("\nConservation Check:\n"); conservation(); }

  free_vec(S.w); free_vec(S.ytmp); free_vec(S.yerr);
  for (int i = 0; i < S.T.s; i++) free_vec(S.k[i]);
  eulerb200_destroy(ctx);
  return 0;
}
