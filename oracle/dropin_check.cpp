/* ---------------------------------------------------------------------------
 * dropin_check.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Links, into one executable,
 *   (1) the reference's utilities.cpp compiled with -DfEuler=ref_fEuler
 *       -Dstability=ref_stability (the UNMODIFIED reference implementation), and
 *   (2) sundials-manyvector-demo_b200/host/feuler_dropin.cpp (our fEuler / stability with
 *       the reference's signatures, on top of libeulerb200.so),
 * builds ONE EulerData object with the reference's own SetupDecomp and ONE set of
 * N_Vectors (5 fluid + 1 chemistry sub-vectors in an MPIManyVector), and calls both
 * callbacks on them: the drop-in claim of SURVEY.md section 8(b), checked end to end.
 * Built by `make dropin` into oracle/_ref/dropin_check_nvar<N>; needs a GPU to run.
 * Prints one line per case and "DROPIN_CHECK PASS" / "DROPIN_CHECK FAIL".
 * ------------------------------------------------------------------------- */
#include <euler3D.hpp>
#include <cstdint>
#include <cmath>
#include <cstdlib>

int ref_fEuler(realtype t, N_Vector w, N_Vector wdot, void* user_data);
int ref_stability(N_Vector w, realtype t, realtype* dt_stab, void* user_data);
extern "C" void eulerb200_dropin_release(void* user_data);
void shim_set_world(int nprocs);
void shim_set_rank(int rank);

static double g_gmy = 0.0;
static int g_varying = 0;       // EB_DROPIN_VARYING=1: a hook that depends on position and time
int external_forces(const realtype& t, N_Vector G, const EulerData& udata)
{
  if (g_varying) {
    const long N = udata.nxl * udata.nyl * udata.nzl;
    const int fields[3] = {1, 2, 4};
    for (int q = 0; q < 3; q++) {
      realtype* g = N_VGetSubvectorArrayPointer_MPIManyVector(G, fields[q]);
      for (long i = 0; i < N; i++) g[i] = 0.05 * std::sin(0.1 * i + fields[q]) + 0.01 * t;
    }
    if (udata.nchem > 0) {
      realtype* g = N_VGetSubvectorArrayPointer_MPIManyVector(G, 5);
      for (long i = 0; i < N * udata.nchem; i++) g[i] = 1e-3 * std::cos(0.3 * i) * (1.0 + t);
    }
    return 0;
  }
  if (g_gmy == 0.0) return 0;
  realtype* g = N_VGetSubvectorArrayPointer_MPIManyVector(G, 2);
  for (long i = 0; i < udata.nxl * udata.nyl * udata.nzl; i++) g[i] = g_gmy;
  return 0;
}

static uint64_t rng_state = 88172645463325252ull;
static double urand()
{
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}

struct Vec { N_Vector sub[6]; N_Vector v; int nsub; };
static void make_vec(Vec& x, EulerData& u)
{
  const long N = u.nxl * u.nyl * u.nzl;
  x.nsub = 5 + (u.nchem > 0 ? 1 : 0);
  for (int f = 0; f < 5; f++) x.sub[f] = N_VNew_Serial(N, u.ctx);
  if (u.nchem > 0) x.sub[5] = N_VNew_Serial(N * u.nchem, u.ctx);
  x.v = N_VMake_MPIManyVector(u.comm, x.nsub, x.sub, u.ctx);
}

static int run_case(long nx, long ny, long nz, const int bc[6], double gmy)
{
  EulerData u;
  u.nx = nx; u.ny = ny; u.nz = nz;
  u.xlbc = bc[0]; u.xrbc = bc[1]; u.ylbc = bc[2]; u.yrbc = bc[3]; u.zlbc = bc[4]; u.zrbc = bc[5];
  u.gamma = 1.4; u.cfl = 0.5;
  g_gmy = gmy;
  if (u.SetupDecomp() != 0) return 1;
  Vec w, a, b;
  make_vec(w, u); make_vec(a, u); make_vec(b, u);
  const long N = u.nxl * u.nyl * u.nzl;
  for (long i = 0; i < N; i++) {
    const double rho = 1 + 0.5 * urand(), vx = 0.3 * (urand() - 0.5), vy = 0.3 * (urand() - 0.5),
                 vz = 0.3 * (urand() - 0.5), p = 1 + 0.5 * urand();
    w.sub[0]->data[i] = rho; w.sub[1]->data[i] = rho * vx; w.sub[2]->data[i] = rho * vy; w.sub[3]->data[i] = rho * vz;
    w.sub[4]->data[i] = u.eos_inv(rho, rho * vx, rho * vy, rho * vz, p);
    for (int v = 0; v < u.nchem; v++) w.sub[5]->data[i * u.nchem + v] = urand();
  }
  const double tcall = g_varying ? 0.3 : 0.0;
  const int r1 = ref_fEuler(tcall, w.v, a.v, (void*)&u);
  const int r2 = fEuler(tcall, w.v, b.v, (void*)&u);
  double worst = 0.0;
  double mom = 0.0;
  for (int f = 1; f <= 3; f++) for (long i = 0; i < N; i++) mom = std::max(mom, std::fabs(a.sub[f]->data[i]));
  for (int f = 0; f < w.nsub; f++) {
    double scale = 0.0, err = 0.0;
    for (long i = 0; i < a.sub[f]->length; i++) {
      scale = std::max(scale, std::fabs(a.sub[f]->data[i]));
      err = std::max(err, std::fabs(a.sub[f]->data[i] - b.sub[f]->data[i]));
    }
    if (f >= 1 && f <= 3) scale = mom;
    worst = std::max(worst, scale > 0 ? err / scale : err);
  }
  realtype dt1 = 0, dt2 = 0;
  const int s1 = ref_stability(w.v, 0.0, &dt1, (void*)&u);
  const int s2 = stability(w.v, 0.0, &dt2, (void*)&u);
  const double dterr = std::fabs(dt1 - dt2) / dt1;
  const bool ok = r1 == 0 && r2 == 0 && s1 == 0 && s2 == 0 && worst <= 1e-12 && dterr <= 1e-14;
  if (g_varying) printf("[position- and time-dependent external_forces hook, run before every evaluation] ");
  printf("NVAR=%d grid %ldx%ldx%ld bc [%d %d %d %d %d %d] Gmy=%g : ret %d/%d  max normwise err %.3e  dt_stab rel err %.1e  %s\n",
         NVAR, nx, ny, nz, bc[0], bc[1], bc[2], bc[3], bc[4], bc[5], gmy, r1, r2, worst, dterr, ok ? "ok" : "MISMATCH");
  eulerb200_dropin_release((void*)&u);
  return ok ? 0 : 1;
}

int main()
{
  shim_set_world(1); shim_set_rank(0);
  int bad = 0;
  g_varying = (getenv("EB_DROPIN_VARYING") != NULL);
  const int per[6] = {0, 0, 0, 0, 0, 0}, neu[6] = {1, 1, 1, 1, 1, 1}, rt[6] = {0, 0, 3, 3, 1, 1}, refl[6] = {3, 3, 3, 3, 3, 3};
  bad += run_case(24, 20, 16, per, 0.0);
  bad += run_case(24, 20, 16, neu, 0.0);
  bad += run_case(16, 24, 3, rt, -0.1);
  bad += run_case(3, 32, 24, neu, 0.0);
  bad += run_case(40, 3, 3, refl, 0.0);
  printf(bad ? "DROPIN_CHECK FAIL\n" : "DROPIN_CHECK PASS\n");
  return bad ? 1 : 0;
}
