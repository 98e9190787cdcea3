// ---------------------------------------------------------------------------
// problems.hpp -- problem plug-ins of the native driver: initial and
// analytic states of the reference's explicit test problems and a flat solution file with the
// reference's dataset order.  Plain C++, no CUDA, no dependency on libeulerb200: the driver
// builds states on the host and copies them to the device once, and tests/ compiles this
// header alone to pin it against the golden fixtures.
//
//   sod_{x,y,z}                 src/sod.cpp:50-55,120-160 (state), :214-379 (exact solution)
//   linear_advection_{x,y,z}    src/linear_advection.cpp:51-68,117-131
//   rayleigh_taylor             src/rayleigh_taylor.cpp:48-53,104-109
//   hurricane_{xy,yz,zx}        src/hurricane.cpp:58-60,120-183 (+ colour-stripe tracers)
//   fluid_blast, primordial_blast   src/fluid_blast.cpp:65-267, src/primordial_blast.cpp:64-309
//                               (fluid state + the ten passive tracers; no chemistry network)
//   solution files              src/io.cpp:716-930 (output_solution), :940-1150 (read_restart)
// ---------------------------------------------------------------------------
#ifndef EULERB200_PROBLEMS_HPP
#define EULERB200_PROBLEMS_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace eb_problems {

const double PI = 3.14159265358979323846;

struct Problem {
  std::string name;
  long nx, ny, nz;
  double xl, xr, yl, yr, zl, zr, gamma;
  int nchem = 0, nprocs = 1;
  double MassUnits = 1.0, LengthUnits = 1.0, TimeUnits = 1.0;
  // euler3D.hpp:385-393
  double DensityUnits() const { return MassUnits / LengthUnits / LengthUnits / LengthUnits; }
  double MomentumUnits() const { return MassUnits / LengthUnits / LengthUnits / TimeUnits; }
  double EnergyUnits() const { return MassUnits / LengthUnits / TimeUnits / TimeUnits; }
  bool is_blast() const { return name == "fluid_blast" || name == "primordial_blast"; }
  double dx() const { return (xr - xl) / nx; }
  double dy() const { return (yr - yl) / ny; }
  double dz() const { return (zr - zl) / nz; }
  char axis() const { return name[name.size() - 1]; }
};

// exact Riemann solution of the Sod tube (sod.cpp:214-379; pL > pR branch)
inline double fsecant(double p4, double p1, double p5, double rho1, double rho5, double g)
{
  const double z = p4 / p5 - 1.0, c1 = sqrt(g * p1 / rho1), c5 = sqrt(g * p5 / rho5);
  const double fact = (g - 1.0) / (2 * g) * (c5 / c1) * z / sqrt(1.0 + (g + 1.0) / (2 * g) * z);
  return p1 * pow(1.0 - fact, 2 * g / (g - 1.0)) - p4;
}
inline void exact_riemann(double t, double x, double xI, double g, double& rho, double& u, double& p)
{
  const double rho1 = 1.0, p1 = 1.0, rho5 = 0.125, p5 = 0.1;
  double p40 = p1, p41 = p5, f0 = fsecant(p40, p1, p5, rho1, rho5, g), p4 = p41;
  for (int it = 0; it < 50; it++) {
    const double f1 = fsecant(p41, p1, p5, rho1, rho5, g);
    if (f1 == f0) break;
    p4 = p41 - (p41 - p40) * f1 / (f1 - f0);
    if (fabs(p4 - p41) / fabs(p41) < 1e-14) break;
    p40 = p41; p41 = p4; f0 = f1;
  }
  const double z = p4 / p5 - 1.0, c5 = sqrt(g * p5 / rho5), gm1 = g - 1.0, gp1 = g + 1.0;
  const double fact = sqrt(1.0 + 0.5 * gp1 * z / g);
  const double u4 = c5 * z / (g * fact), rho4 = rho5 * (1.0 + 0.5 * gp1 * z / g) / (1.0 + 0.5 * gm1 * z / g);
  const double w = c5 * fact, p3 = p4, u3 = u4, rho3 = rho1 * pow(p3 / p1, 1.0 / g);
  const double c1 = sqrt(g * p1 / rho1), c3 = sqrt(g * p3 / rho3);
  const double xsh = xI + w * t, xcd = xI + u3 * t, xft = xI + (u3 - c3) * t, xhd = xI - c1 * t;
  if (x < xhd) { rho = rho1; p = p1; u = 0.0; }
  else if (x < xft) {
    u = 2.0 / gp1 * (c1 + (x - xI) / t);
    const double f = 1.0 - 0.5 * gm1 * u / c1;
    rho = rho1 * pow(f, 2.0 / gm1); p = p1 * pow(f, 2.0 * g / gm1);
  }
  else if (x < xcd) { rho = rho3; p = p3; u = u3; }
  else if (x < xsh) { rho = rho4; p = p4; u = u4; }
  else { rho = rho5; p = p5; u = 0.0; }
}

// Critical-rotation solution of the hurricane problems (hurricane.cpp:236-303): density and the
// three momenta of cell (i,j,k) at time t (the reference's diagnostics compare these four fields).
// In the rotation plane (a,b): inside r < 2 t sqrt(p0') the gas has expanded into a paraboloid,
// rho = r^2 / (8 A t^2) with m = rho ((a+b), (b-a)) / (2t); outside it still has rho0 and the
// velocity of a fluid parcel that started on the v0 circle.  p0' = A gamma rho0^(gamma-1).
inline void hurricane_true(const Problem& P, double t, long i, long j, long k, double w4[4])
{
  const double x = (i + 0.5) * P.dx() + P.xl, y = (j + 0.5) * P.dy() + P.yl, z = (k + 0.5) * P.dz() + P.zl;
  const std::string pl = P.name.substr(P.name.size() - 2);
  const double a = pl == "xy" ? x : (pl == "zx" ? z : y), b = pl == "xy" ? y : (pl == "zx" ? x : z);
  const double A = 25.0, rho0 = 1.0;
  const double p0p = A * P.gamma * pow(rho0, P.gamma - 1.0);
  double r = sqrt(a * a + b * b);
  if (r == 0.0) r = 1e-14;
  const double ca = a / r, sb = b / r;
  double rho, ma, mb;
  if (r < 2.0 * t * sqrt(p0p)) {
    rho = r * r / (8.0 * A * t * t);
    ma = rho * (a + b) / (2.0 * t);
    mb = rho * (b - a) / (2.0 * t);
  } else {
    const double swirl = sqrt(2.0 * p0p) * sqrt(r * r - 2.0 * t * t * p0p);
    rho = rho0;
    ma = rho0 * (2.0 * t * p0p * ca + swirl * sb) / r;
    mb = rho0 * (2.0 * t * p0p * sb - swirl * ca) / r;
  }
  w4[0] = rho; w4[1] = w4[2] = w4[3] = 0.0;
  if (pl == "xy") { w4[1] = ma; w4[2] = mb; } else if (pl == "zx") { w4[3] = ma; w4[1] = mb; } else { w4[2] = ma; w4[3] = mb; }
}

// Analytic / initial state of cell (i,j,k) at time t; returns false if the problem has no
// analytic solution for t > 0 (then only t = t0 is meaningful).
inline bool state_at(const Problem& P, double t, long i, long j, long k, double w[5])
{
  const double x = (i + 0.5) * P.dx() + P.xl, y = (j + 0.5) * P.dy() + P.yl, z = (k + 0.5) * P.dz() + P.zl;
  double rho = 1.0, m[3] = {0, 0, 0}, p = 1.0;
  bool analytic = true;
  if (P.name.compare(0, 3, "sod") == 0) {
    const int a = P.axis() - 'x';
    const double s = a == 0 ? x : (a == 1 ? y : z);
    double u = 0.0;
    if (t > 0.0) exact_riemann(t, s, 0.5, P.gamma, rho, u, p);
    else { rho = s < 0.5 ? 1.0 : 0.125; p = s < 0.5 ? 1.0 : 0.1; }
    m[a] = rho * u;
  } else if (P.name.compare(0, 16, "linear_advection") == 0) {
    const int a = P.axis() - 'x';
    const double s = a == 0 ? x : (a == 1 ? y : z);
    rho = 1.0 + 0.1 * sin(2.0 * PI * (s - 0.5 * t));
    m[a] = 0.5 * rho;
  } else if (P.name == "rayleigh_taylor") {
    rho = y > 0.0 ? 2.0 : 1.0;
    m[1] = rho * 0.01 * (1.0 + cos(4.0 * PI * x)) * (1.0 + cos(3.0 * PI * y));
    p = 2.5 - 0.1 * rho * y;
    analytic = false;
  } else if (P.name.compare(0, 9, "hurricane") == 0) {
    const std::string pl = P.name.substr(P.name.size() - 2);
    const double a = pl == "xy" ? x : (pl == "zx" ? z : y), b = pl == "xy" ? y : (pl == "zx" ? x : z);
    double r = sqrt(a * a + b * b);
    if (r == 0.0) r = 1e-14;
    const double ma = 10.0 * (b / r), mb = -10.0 * (a / r);
    if (pl == "xy") { m[0] = ma; m[1] = mb; } else if (pl == "zx") { m[2] = ma; m[0] = mb; } else { m[1] = ma; m[2] = mb; }
    p = 25.0;
    analytic = false;
  } else {
    fprintf(stderr, "unknown problem '%s'\n", P.name.c_str());
    exit(1);
  }
  w[0] = rho; w[1] = m[0]; w[2] = m[1]; w[3] = m[2];
  w[4] = p / (P.gamma - 1.0) + (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) * 0.5 / rho;      // eos_inv
  return analytic;
}

// Tracers of cell (i,j,k) at t0 for the non-blast problems: hurricane colours the fluid in
// nchem angular stripes (hurricane.cpp:122-125,172-180); every other problem starts them at zero.
inline void tracers_at(const Problem& P, long i, long j, long k, double* c)
{
  for (int v = 0; v < P.nchem; v++) c[v] = 0.0;
  if (P.name.compare(0, 9, "hurricane") != 0) return;
  const double x = (i + 0.5) * P.dx() + P.xl, y = (j + 0.5) * P.dy() + P.yl, z = (k + 0.5) * P.dz() + P.zl;
  const std::string pl = P.name.substr(P.name.size() - 2);
  const double theta = pl == "xy" ? atan2(y, x) : (pl == "zx" ? atan2(x, z) : atan2(z, y));
  for (int v = 0; v < P.nchem; v++) {
    const double lo = -PI + v * 2.0 * PI / P.nchem, hi = -PI + (v + 1) * 2.0 * PI / P.nchem;
    c[v] = (theta >= lo && theta < hi) ? 1.0 : 0.0;
  }
}

// 10*nprocs Gaussian clumps drawn from std::mt19937_64 seeded with the rank count
// (fluid_blast.cpp:98-128, primordial_blast.cpp:102-123): centre, radius in cells, strength.
struct Clump { double cx, cy, cz, cr, cs; };
inline std::vector<Clump> blast_clumps(const Problem& P, double max_strength)
{
  std::mt19937_64 gen(P.nprocs);
  std::uniform_real_distribution<double> cx_d(P.xl, P.xr), cy_d(P.yl, P.yr), cz_d(P.zl, P.zr), cr_d(3.0, 6.0),
      cs_d(0.0, max_strength);
  std::vector<Clump> out(10 * (size_t)P.nprocs);
  for (auto& c : out) { c.cx = cx_d(gen); c.cy = cy_d(gen); c.cz = cz_d(gen); c.cr = cr_d(gen); c.cs = cs_d(gen); }
  return out;
}

// Clumpy neutral primordial gas at rest plus a hot dense central clump, in code units
// (fluid_blast.cpp:140-264, primordial_blast.cpp:180-297).  fluid[f] are the five SoA fields of
// length nx*ny*nz; chem (may be NULL when nchem == 0) is the AoS tracer block: eight number
// densities, electron density, gas energy.
inline int blast_state(const Problem& P, double* const fluid[5], double* chem)
{
  if (P.nchem != 0 && P.nchem != 10) { fprintf(stderr, "the blast problems carry 0 or 10 species\n"); return -1; }
  const double mH = 1.67e-24, kboltz = 1.3806488e-16, Hfrac = 0.76, m_amu = 1.66053904e-24;
  const double density0 = 1e2 * mH, tiny = 1e-40, small = 1e-12;
  const bool fl = P.name == "fluid_blast";          // the two files differ in three constants
  const double max_strength = fl ? 10.0 : 5.0, blast_density = fl ? 10.0 : 5.0;
  const double blast_temp = fl ? 10.0 * 5.0 : 10.0 * (5.0 - 1.0);
  const std::vector<Clump> clumps = blast_clumps(P, max_strength);
  const double bx = P.xl + 0.5 * (P.xr - P.xl), by = P.yl + 0.5 * (P.yr - P.yl), bz = P.zl + 0.5 * (P.zr - P.zl);
  const double br = 0.1 * std::min(P.xr - P.xl, std::min(P.yr - P.yl, P.zr - P.zl));
  const double wH = 1.00794 * mH, wHe = 4.002602 * mH;
  for (long k = 0; k < P.nz; k++)
    for (long j = 0; j < P.ny; j++)
      for (long i = 0; i < P.nx; i++) {
        const double x = (i + 0.5) * P.dx() + P.xl, y = (j + 0.5) * P.dy() + P.yl, z = (k + 0.5) * P.dz() + P.zl;
        double density = 1.0;
        for (const Clump& c : clumps) {
          const double cr = c.cr * P.dx();
          const double rsq = (x - c.cx) * (x - c.cx) + (y - c.cy) * (y - c.cy) + (z - c.cz) * (z - c.cz);
          density += c.cs * exp(-2.0 * rsq / cr / cr);
        }
        density *= density0;
        const double rsq = (x - bx) * (x - bx) + (y - by) * (y - by) + (z - bz) * (z - bz);
        const double bump = exp(-2.0 * rsq / br / br);
        density += density0 * blast_density * bump;
        const double T = 10.0 + blast_temp * bump;
        const bool inside = rsq / br / br < 2.0;
        const double lo = 1.0e-3 * density;
        const double H2I = inside ? tiny * density : lo, H2II = H2I, HM = H2I;
        const double HII = inside ? small * density : lo, HeII = HII, HeIII = HII;
        const double HeI = (1.0 - Hfrac) * density - HeII - HeIII;
        const double HI = density - (H2I + H2II + HII + HM + HeI + HeII + HeIII);
        const double nH2I = H2I / (2 * wH), nH2II = H2II / (2 * wH), nHII = HII / wH, nHM = HM / wH;
        const double nHeII = HeII / wHe, nHeIII = HeIII / wHe, nHeI = HeI / wHe, nHI = HI / wH;
        const double ndens = nH2I + nH2II + nHII + nHM + nHeII + nHeIII + nHeI + nHI;
        const double ge = (kboltz * T * ndens) / (density * (P.gamma - 1.0));
        const long c = i + P.nx * (j + P.ny * k);
        fluid[0][c] = density / P.DensityUnits();
        fluid[1][c] = fluid[2][c] = fluid[3][c] = 0.0;
        fluid[4][c] = ge / P.EnergyUnits();
        if (P.nchem == 10) {
          const double de = (nHII + nHeII + 2 * nHeIII - nHM + nH2II) * mH;
          const double sp[10] = {nH2I, nH2II, nHI, nHII, nHM, nHeI, nHeII, nHeIII, de / m_amu, ge};
          for (int v = 0; v < 10; v++) chem[10 * c + v] = sp[v];
        }
      }
  return 0;
}

// initial_conditions(t0, w, udata) of whichever problem file the reference was linked with
inline int initial_conditions(const Problem& P, double t0, double* const fluid[5], double* chem, bool* analytic)
{
  if (P.is_blast()) { *analytic = false; return blast_state(P, fluid, chem); }
  if (P.name.compare(0, 3, "sod") == 0) t0 = 0.0;      // sod.cpp:120-160 sets the jump whatever t is
  for (long k = 0; k < P.nz; k++)
    for (long j = 0; j < P.ny; j++)
      for (long i = 0; i < P.nx; i++) {
        double w5[5];
        *analytic = state_at(P, t0, i, j, k, w5);
        const long c = i + P.nx * (j + P.ny * k);
        for (int f = 0; f < 5; f++) fluid[f][c] = w5[f];
        if (P.nchem > 0) tracers_at(P, i, j, k, chem + P.nchem * c);
      }
  return 0;
}

// ---- solution files ---------------------------------------------------------------------
// output-<iout>.eb200: what output_solution (io.cpp:716-930) stores in output-<iout>.hdf5, as one
// flat little-endian file (there is no HDF5 here):
//   char[8] "EB200OUT" | int32 version = 1 | int32 nchem | int64 nx, ny, nz | double time |
//   double domain[6] = zl, zr, yl, yr, xl, xr (the reference's order, io.cpp:827-829) |
//   datasets of nx*ny*nz doubles, x fastest, in the reference's order: Density, x-Momentum,
//   y-Momentum, z-Momentum, TotalEnergy (scaled to CGS with the unit factors, io.cpp:887-891),
//   Chemical-000 ... (one contiguous dataset per species, io.cpp:910-925).
const int SOLUTION_HEADER_BYTES = 8 + 4 + 4 + 3 * 8 + 8 + 6 * 8;

inline std::string solution_name(int iout)
{
  char nm[64];
  snprintf(nm, sizeof nm, "output-%07i.eb200", iout);
  return nm;
}

inline int write_solution(const std::string& path, const Problem& P, double t, const double* const fluid[5],
                          const double* chem)
{
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return -1;
  const int32_t head[2] = {1, (int32_t)P.nchem};
  const int64_t n[3] = {P.nx, P.ny, P.nz};
  const double dom[6] = {P.zl, P.zr, P.yl, P.yr, P.xl, P.xr};
  const size_t N = (size_t)(P.nx * P.ny * P.nz);
  bool ok = fwrite("EB200OUT", 1, 8, fp) == 8 && fwrite(head, 4, 2, fp) == 2 && fwrite(n, 8, 3, fp) == 3 &&
            fwrite(&t, 8, 1, fp) == 1 && fwrite(dom, 8, 6, fp) == 6;
  const double scale[5] = {P.DensityUnits(), P.MomentumUnits(), P.MomentumUnits(), P.MomentumUnits(), P.EnergyUnits()};
  std::vector<double> tmp(N);
  for (int f = 0; f < 5 && ok; f++) {
    for (size_t c = 0; c < N; c++) tmp[c] = scale[f] * fluid[f][c];
    ok = fwrite(tmp.data(), 8, N, fp) == N;
  }
  for (int v = 0; v < P.nchem && ok; v++) {
    for (size_t c = 0; c < N; c++) tmp[c] = chem[c * P.nchem + v];
    ok = fwrite(tmp.data(), 8, N, fp) == N;
  }
  return (fclose(fp) == 0 && ok) ? 0 : -1;
}

// read_restart (io.cpp:940-1150): the file must describe the same grid, species count and domain
inline int read_solution(const std::string& path, const Problem& P, double* t, double* const fluid[5], double* chem)
{
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { fprintf(stderr, "read_solution: cannot open %s\n", path.c_str()); return -1; }
  char magic[8];
  int32_t head[2];
  int64_t n[3];
  double dom[6];
  bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "EB200OUT", 8) == 0 && fread(head, 4, 2, fp) == 2 &&
            fread(n, 8, 3, fp) == 3 && fread(t, 8, 1, fp) == 1 && fread(dom, 8, 6, fp) == 6;
  if (ok && (head[0] != 1 || head[1] != P.nchem || n[0] != P.nx || n[1] != P.ny || n[2] != P.nz)) {
    fprintf(stderr, "read_solution: %s holds a %lld x %lld x %lld grid with %d species, the run asks for "
            "%ld x %ld x %ld with %d\n", path.c_str(), (long long)n[0], (long long)n[1], (long long)n[2], (int)head[1],
            P.nx, P.ny, P.nz, P.nchem);
    ok = false;
  }
  const size_t N = (size_t)(P.nx * P.ny * P.nz);
  const double scale[5] = {P.DensityUnits(), P.MomentumUnits(), P.MomentumUnits(), P.MomentumUnits(), P.EnergyUnits()};
  std::vector<double> tmp(N);
  for (int f = 0; f < 5 && ok; f++) {
    ok = fread(tmp.data(), 8, N, fp) == N;
    for (size_t c = 0; c < N && ok; c++) fluid[f][c] = tmp[c] / scale[f];
  }
  for (int v = 0; v < P.nchem && ok; v++) {
    ok = fread(tmp.data(), 8, N, fp) == N;
    for (size_t c = 0; c < N && ok; c++) chem[c * P.nchem + v] = tmp[c];
  }
  fclose(fp);
  return ok ? 0 : -1;
}

}  // namespace eb_problems
#endif
