/* ---------------------------------------------------------------------------
 * ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A C-callable harness around the UNMODIFIED reference fluid RHS.  It is compiled
 * together with /root/reference/src/utilities.cpp (read from where it lies; never
 * copied into this repository) against oracle/shim/ into oracle/_ref/libref_nvar<N>.so,
 * one library per NVAR (the reference fixes nchem = NVAR-5 at compile time,
 * euler3D.hpp:52-59,297).
 *
 * What it drives (all reference code, none of ours):
 *   fEuler            utilities.cpp:17-253
 *   face_flux         utilities.cpp:270-479
 *   stability         utilities.cpp:483-528
 *   EulerData::SetupDecomp / ExchangeStart / ExchangeEnd   euler3D.hpp:396-1191
 *
 * The link-time hook external_forces (euler3D.hpp:1454) is provided here as
 * "assign a constant per fluid field", which covers every shipped problem:
 * zero everywhere except Rayleigh-Taylor's Gmy = -0.1 (rayleigh_taylor.cpp:117-128).
 *
 * Grids are passed GLOBAL; with nprocs > 1 the state is scattered over virtual
 * ranks (threads, see shim_mpi.cpp) using the reference's own SetupDecomp.
 * ------------------------------------------------------------------------- */
#include <euler3D.hpp>
#include <thread>
#include <atomic>

void shim_set_world(int nprocs);
void shim_set_rank(int rank);

static double g_forcing[5] = {0.0, 0.0, 0.0, 0.0, 0.0};

int external_forces(const realtype& t, N_Vector G, const EulerData& udata)
{
  (void)t;
  const long int N = udata.nxl * udata.nyl * udata.nzl;
  for (int f = 0; f < 5; f++) {
    if (g_forcing[f] == 0.0) continue;   /* wdot was already zeroed (utilities.cpp:28) */
    realtype* g = N_VGetSubvectorArrayPointer_MPIManyVector(G, f);
    if (g == NULL) return -1;
    for (long int i = 0; i < N; i++) g[i] = g_forcing[f];
  }
  return 0;
}

namespace {

struct Problem {
  long n[3];
  double box[6];
  int bc[6];
  double gamma, cfl;
};

void configure(EulerData& u, const Problem& p)
{
  u.nx = p.n[0]; u.ny = p.n[1]; u.nz = p.n[2];
  u.xl = p.box[0]; u.xr = p.box[1]; u.yl = p.box[2]; u.yr = p.box[3]; u.zl = p.box[4]; u.zr = p.box[5];
  u.xlbc = p.bc[0]; u.xrbc = p.bc[1]; u.ylbc = p.bc[2]; u.yrbc = p.bc[3]; u.zlbc = p.bc[4]; u.zrbc = p.bc[5];
  u.gamma = p.gamma; u.cfl = p.cfl;
}

struct LocalState {
  N_Vector sub[6];
  N_Vector w;
  int nsub;
};

void make_state(LocalState& s, const EulerData& u)
{
  const long N = u.nxl * u.nyl * u.nzl;
  s.nsub = 5 + (u.nchem > 0 ? 1 : 0);
  for (int f = 0; f < 5; f++) s.sub[f] = N_VNew_Serial(N, u.ctx);
  if (u.nchem > 0) s.sub[5] = N_VNew_Serial(N * u.nchem, u.ctx);
  s.w = N_VMake_MPIManyVector(u.comm, s.nsub, s.sub, u.ctx);
}
void free_state(LocalState& s)
{
  for (int f = 0; f < s.nsub; f++) N_VDestroy(s.sub[f]);
  N_VDestroy(s.w);
}

void scatter(LocalState& s, const EulerData& u, const double* const* g)
{
  for (long k = 0; k < u.nzl; k++)
    for (long j = 0; j < u.nyl; j++)
      for (long i = 0; i < u.nxl; i++) {
        const long l = INDX(i, j, k, u.nxl, u.nyl, u.nzl);
        const long G = INDX(i + u.is, j + u.js, k + u.ks, u.nx, u.ny, u.nz);
        for (int f = 0; f < 5; f++) s.sub[f]->data[l] = g[f][G];
        for (int v = 0; v < u.nchem; v++) s.sub[5]->data[v + u.nchem * l] = g[5][v + u.nchem * G];
      }
}
void gather(const LocalState& s, const EulerData& u, double* const* g)
{
  for (long k = 0; k < u.nzl; k++)
    for (long j = 0; j < u.nyl; j++)
      for (long i = 0; i < u.nxl; i++) {
        const long l = INDX(i, j, k, u.nxl, u.nyl, u.nzl);
        const long G = INDX(i + u.is, j + u.js, k + u.ks, u.nx, u.ny, u.nz);
        for (int f = 0; f < 5; f++) g[f][G] = s.sub[f]->data[l];
        for (int v = 0; v < u.nchem; v++) g[5][v + u.nchem * G] = s.sub[5]->data[v + u.nchem * l];
      }
}

template <class F> void run_ranks(int nprocs, F body)
{
  shim_set_world(nprocs);
  if (nprocs == 1) { shim_set_rank(0); body(0); return; }
  std::vector<std::thread> th;
  for (int r = 0; r < nprocs; r++) th.emplace_back([=] { shim_set_rank(r); body(r); });
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int refdrv_nvar(void) { return NVAR; }

void refdrv_set_forcing(const double* g5) { for (int f = 0; f < 5; f++) g_forcing[f] = g5[f]; }

/* One reference face flux (utilities.cpp:270): w1d is [6][NVAR] row-major. */
void refdrv_face_flux(const double* w1d_in, int idir, double gamma, double* f_face)
{
  shim_set_world(1); shim_set_rank(0);
  EulerData u;
  u.gamma = gamma;
  realtype w1d[6][NVAR];
  for (int l = 0; l < 6; l++) for (int v = 0; v < NVAR; v++) w1d[l][v] = w1d_in[l * NVAR + v];
  face_flux(w1d, idir, f_face, u);
}

/* fEuler on the global grid, decomposed over nprocs virtual ranks.
 * w[0..4] fluid (i + nx*(j + ny*k)), w[5] chem (v + nchem*cell) or NULL; same for wdot.
 * nrep >= 1 evaluations are made; seconds[r] receives the max-over-ranks wall time of
 * evaluation r.  decomp (may be NULL) receives npx,npy,npz.
 * Returns the minimum fEuler return value over ranks and repetitions. */
int refdrv_feuler(int nprocs, const long* n, const double* box, const int* bc, double gamma,
                  double t, const double* const* w, double* const* wdot,
                  int nrep, double* seconds, int* decomp)
{
  Problem p; for (int d = 0; d < 3; d++) p.n[d] = n[d];
  for (int d = 0; d < 6; d++) { p.box[d] = box[d]; p.bc[d] = bc[d]; }
  p.gamma = gamma; p.cfl = 0.0;
  std::atomic<int> rv(0);
  std::vector<std::vector<double> > secs(nprocs, std::vector<double>(nrep, 0.0));
  run_ranks(nprocs, [&](int r) {
    EulerData u; configure(u, p);
    if (u.SetupDecomp() != 0) { rv = -2; return; }
    if (decomp && r == 0) { decomp[0] = u.npx; decomp[1] = u.npy; decomp[2] = u.npz; }
    LocalState s, sd; make_state(s, u); make_state(sd, u);
    scatter(s, u, w);
    for (int it = 0; it < nrep; it++) {
      if (nprocs > 1) MPI_Barrier(u.comm);
      double t0 = MPI_Wtime();
      int ret = fEuler(t, s.w, sd.w, (void*)&u);
      secs[r][it] = MPI_Wtime() - t0;
      if (ret < rv) rv = ret;
      /* the reference leaves PR_RHSEULER running when it bails out early */
      if (ret != 0) { u.profile[PR_RHSEULER].stop(); }
    }
    gather(sd, u, wdot);
    free_state(s); free_state(sd);
  });
  if (seconds) for (int it = 0; it < nrep; it++) {
    double m = 0; for (int r = 0; r < nprocs; r++) m = std::max(m, secs[r][it]);
    seconds[it] = m;
  }
  return rv;
}

/* stability (utilities.cpp:483) on the global grid over nprocs virtual ranks. */
int refdrv_stability(int nprocs, const long* n, const double* box, const int* bc, double gamma,
                     double cfl, const double* const* w, double* dt_out)
{
  Problem p; for (int d = 0; d < 3; d++) p.n[d] = n[d];
  for (int d = 0; d < 6; d++) { p.box[d] = box[d]; p.bc[d] = bc[d]; }
  p.gamma = gamma; p.cfl = cfl;
  std::atomic<int> rv(0);
  run_ranks(nprocs, [&](int r) {
    EulerData u; configure(u, p);
    if (u.SetupDecomp() != 0) { rv = -2; return; }
    LocalState s; make_state(s, u); scatter(s, u, w);
    double dt = 0;
    int ret = stability(s.w, 0.0, &dt, (void*)&u);
    if (ret < rv) rv = ret;
    if (r == 0) *dt_out = dt;
    free_state(s);
  });
  return rv;
}

/* ExchangeStart + ExchangeEnd (euler3D.hpp:577-1191); returns, for virtual rank `rank`,
 * its extents ext = {is,ie,js,je,ks,ke}, neighbours nbr = {ipW,ipE,ipS,ipN,ipB,ipF}
 * (-2 = MPI_PROC_NULL) and copies of its six receive buffers (W,E,S,N,B,F; caller
 * allocates NVAR*3*area doubles each; any may be NULL). */
int refdrv_exchange(int nprocs, const long* n, const int* bc, const double* const* w,
                    int rank, long* ext, int* nbr, double* const* recv)
{
  Problem p; for (int d = 0; d < 3; d++) p.n[d] = n[d];
  for (int d = 0; d < 6; d++) { p.box[d] = (d % 2) ? 1.0 : 0.0; p.bc[d] = bc[d]; }
  p.gamma = 1.4; p.cfl = 0.0;
  std::atomic<int> rv(0);
  run_ranks(nprocs, [&](int r) {
    EulerData u; configure(u, p);
    if (u.SetupDecomp() != 0) { rv = -2; return; }
    LocalState s; make_state(s, u); scatter(s, u, w);
    if (u.ExchangeStart(s.w) != 0) rv = -1;
    if (u.ExchangeEnd() != 0) rv = -1;
    if (r == rank) {
      ext[0] = u.is; ext[1] = u.ie; ext[2] = u.js; ext[3] = u.je; ext[4] = u.ks; ext[5] = u.ke;
      nbr[0] = u.ipW; nbr[1] = u.ipE; nbr[2] = u.ipS; nbr[3] = u.ipN; nbr[4] = u.ipB; nbr[5] = u.ipF;
      const long sz[3] = {(NVAR) * 3 * u.nyl * u.nzl, (NVAR) * u.nxl * 3 * u.nzl, (NVAR) * u.nxl * u.nyl * 3};
      const realtype* src[6] = {u.Wrecv, u.Erecv, u.Srecv, u.Nrecv, u.Brecv, u.Frecv};
      for (int f = 0; f < 6; f++) if (recv[f]) memcpy(recv[f], src[f], sizeof(double) * sz[f / 2]);
    }
    free_state(s);
  });
  return rv;
}

}  // extern "C"
