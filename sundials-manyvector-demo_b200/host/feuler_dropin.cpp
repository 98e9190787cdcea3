// ---------------------------------------------------------------------------
// feuler_dropin.cpp -- drop-in replacements for the two ARKODE callbacks of the
// reference's fluid path, to be compiled INTO a build of sundials-manyvector-demo in
// place of the definitions in src/utilities.cpp:
//
//     int fEuler(realtype t, N_Vector w, N_Vector wdot, void* user_data);    utilities.cpp:17
//     int stability(N_Vector w, realtype t, realtype* dt_stab, void* user_data);  :483
//
// Same signatures, same EulerData (the reference's own header is included unchanged), same
// 5 fluid + 1 chemistry MPIManyVector composition; the drivers (euler3D_main.cpp:194,273,
// multirate_chem_hydro_main.cpp:1046, imex_chem_hydro_main.cpp:963) call them unchanged.
// Everything numerical happens behind the C ABI of include/eulerb200.h.
//
// What is taken from EulerData: nxl,nyl,nzl, dx,dy,dz, the six BC codes, gamma, nchem,
// myid/nprocs and the six neighbour ranks (euler3D.hpp:198-256).  The reference's host
// scratch (xflux..zflux, the 12 halo buffers) is simply not used.
//
// Forcing: the reference calls the link-time hook external_forces(t, wdot, udata) inside
// fEuler (utilities.cpp:65), which ASSIGNS G into wdot.  Every shipped problem assigns a
// constant per field (zero, or Gmy = -0.1 for Rayleigh-Taylor), so the hook is evaluated
// once on a host probe vector when the context is created (at t0 and at a later time) and,
// if it is such a constant, its values are handed to the kernel.  Any other hook is run on
// wdot before every evaluation exactly as the reference does (utilities.cpp:28,65) and the
// kernel subtracts the flux divergence from what it finds there
// (eulerb200_set_forcing_in_wdot).
//
// Multi-rank: the NCCL id (MPI_Bcast) and, with EULERB200_HALO=p2p, the CUDA-IPC mailbox handles of
// the peer-store halo transport (MPI_Allgather) travel once, at context creation, over the
// reference's own communicator udata->comm -- no MPI call is left on the per-RHS path.
// ---------------------------------------------------------------------------
#include <euler3D.hpp>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "eulerb200.h"

namespace {

struct Binding { eulerb200_ctx* ctx; bool hook_per_call; };
std::map<const EulerData*, Binding>& bindings()
{
  static std::map<const EulerData*, Binding> m;
  return m;
}

// Evaluate the external_forces hook on a probe the size of the local block, at t0 and at a
// later time, and reduce it to one constant per fluid field.  Returns 0 if it is such a
// constant (and zero for the species), 1 if it is not, -1 if the hook fails.
int probe_forcing(EulerData* udata, double forcing[5])
{
  const long N = udata->nxl * udata->nyl * udata->nzl;
  N_Vector sub[6];
  const int nsub = 5 + (udata->nchem > 0 ? 1 : 0);
  for (int f = 0; f < 5; f++) sub[f] = N_VNew_Serial(N, udata->ctx);
  if (udata->nchem > 0) sub[5] = N_VNew_Serial(N * udata->nchem, udata->ctx);
  N_Vector G = N_VMake_MPIManyVector(udata->comm, nsub, sub, udata->ctx);
  int ret = 0;
  const double times[2] = {udata->t0, udata->t0 + 0.37 * (udata->tf - udata->t0) + 1.0};
  for (int pass = 0; pass < 2 && ret == 0; pass++) {
    N_VConst(ZERO, G);
    if (external_forces(times[pass], G, *udata) != 0) { ret = -1; break; }
    for (int f = 0; f < 5 && ret == 0; f++) {
      const realtype* g = N_VGetArrayPointer(sub[f]);
      if (pass == 0) forcing[f] = g[0];
      for (long i = 0; i < N; i++)
        if (g[i] != forcing[f]) { ret = 1; break; }
    }
    if (ret == 0 && udata->nchem > 0) {
      const realtype* g = N_VGetArrayPointer(sub[5]);
      for (long i = 0; i < N * udata->nchem; i++)
        if (g[i] != ZERO) { ret = 1; break; }
    }
  }
  N_VDestroy(G);
  for (int f = 0; f < nsub; f++) N_VDestroy(sub[f]);
  return ret;
}

eulerb200_ctx* context_for(EulerData* udata)
{
  auto it = bindings().find(udata);
  if (it != bindings().end()) return it->second.ctx;
  eulerb200_config c;
  c.nxl = udata->nxl; c.nyl = udata->nyl; c.nzl = udata->nzl;
  c.nchem = udata->nchem;
  c.device = -1;
  c.dx = udata->dx; c.dy = udata->dy; c.dz = udata->dz;
  c.gamma = udata->gamma;
  const int bc[6] = {udata->xlbc, udata->xrbc, udata->ylbc, udata->yrbc, udata->zlbc, udata->zrbc};
  const int nb[6] = {udata->ipW, udata->ipE, udata->ipS, udata->ipN, udata->ipB, udata->ipF};
  for (int f = 0; f < 6; f++) {
    c.bc[f] = bc[f];
    c.nbr[f] = (nb[f] == MPI_PROC_NULL) ? EULERB200_NO_NEIGHBOR : nb[f];
  }
  c.rank = udata->myid;
  c.nranks = udata->nprocs;
  const int probe = probe_forcing(udata, c.forcing);
  if (probe < 0) return NULL;
  if (probe == 1) for (int f = 0; f < 5; f++) c.forcing[f] = 0.0;   // the hook runs before every evaluation
  eulerb200_ctx* ctx = NULL;
  if (eulerb200_create(&c, &ctx) != 0) {
    cerr << "\neulerb200_create failed: " << eulerb200_last_error(NULL) << "\n\n";
    return NULL;
  }
  if (udata->nprocs > 1) {
    char id[EULERB200_UNIQUE_ID_BYTES];
    if (udata->myid == 0 && eulerb200_comm_unique_id(id) != 0) return NULL;
    if (MPI_Bcast(id, EULERB200_UNIQUE_ID_BYTES, MPI_BYTE, 0, udata->comm) != MPI_SUCCESS) return NULL;
    if (eulerb200_comm_attach(ctx, id) != 0) {
      cerr << "\neulerb200_comm_attach failed: " << eulerb200_last_error(ctx) << "\n\n";
      return NULL;
    }
    // Halo transport: NCCL send/recv pairs by default (fastest at 8 GPUs).  With
    // EULERB200_HALO=p2p on every rank the peer-store transport (CUDA IPC over NVLink) is attached
    // instead; it needs every neighbour to be peer-accessible (one NVSwitch node).
    const char* halo = getenv("EULERB200_HALO");
    if (halo != NULL && std::string(halo) == "p2p") {
      std::vector<char> blobs((size_t)EULERB200_P2P_BLOB_BYTES * udata->nprocs);
      char mine[EULERB200_P2P_BLOB_BYTES];
      int ok = eulerb200_p2p_export(ctx, mine) == 0 ? 1 : 0;
      if (MPI_Allgather(mine, EULERB200_P2P_BLOB_BYTES, MPI_BYTE, blobs.data(), EULERB200_P2P_BLOB_BYTES, MPI_BYTE,
                        udata->comm) != MPI_SUCCESS) return NULL;
      double flag = (ok && eulerb200_p2p_attach(ctx, blobs.data()) == 0) ? 1.0 : 0.0;
      if (MPI_Allreduce(MPI_IN_PLACE, &flag, 1, MPI_DOUBLE, MPI_MIN, udata->comm) != MPI_SUCCESS) return NULL;
      if (flag == 0.0) {
        cerr << "\neulerb200: EULERB200_HALO=p2p but the peer-store halo transport is unavailable on some rank "
                "(all GPUs of the run must be peer-accessible); unset it to use NCCL\n\n";
        return NULL;
      }
    }
  }
  if (probe == 1 && eulerb200_set_forcing_in_wdot(ctx, 1) != 0) return NULL;
  bindings()[udata].ctx = ctx;
  bindings()[udata].hook_per_call = (probe == 1);
  return ctx;
}

int subvector_pointers(N_Vector v, int nchem, const char* who, double* out[6])
{
  for (int f = 0; f < 6; f++) out[f] = NULL;
  for (int f = 0; f < 5 + (nchem > 0 ? 1 : 0); f++) {
    out[f] = N_VGetSubvectorArrayPointer_MPIManyVector(v, f);
    if (check_flag((void*)out[f], who, 0)) return -1;
  }
  return 0;
}

}  // namespace

// The device context bound to an EulerData (created on first use): for callers that run their vector
// operations on the device next to the RHS (a CUDA / managed N_Vector build of the driver).
extern "C" eulerb200_ctx* eulerb200_dropin_context(void* user_data)
{
  return context_for((EulerData*)user_data);
}

// Release the device context bound to an EulerData (call next to EulerData::FreeData).
extern "C" void eulerb200_dropin_release(void* user_data)
{
  auto it = bindings().find((const EulerData*)user_data);
  if (it == bindings().end()) return;
  eulerb200_destroy(it->second.ctx);
  bindings().erase(it);
}

int fEuler(realtype t, N_Vector w, N_Vector wdot, void* user_data)
{
  EulerData* udata = (EulerData*)user_data;
  int retval = udata->profile[PR_RHSEULER].start();
  if (check_flag(&retval, "Profile::start (fEuler)", 1)) return -1;

  double *wp[6], *wdp[6];
  if (subvector_pointers(w, udata->nchem, "N_VGetSubvectorArrayPointer (fEuler)", wp)) return -1;
  if (subvector_pointers(wdot, udata->nchem, "N_VGetSubvectorArrayPointer (fEuler)", wdp)) return -1;

  eulerb200_ctx* ctx = context_for(udata);
  if (ctx == NULL) return -1;
  if (bindings()[udata].hook_per_call) {       // utilities.cpp:28,65
    N_VConst(ZERO, wdot);
    retval = external_forces(t, wdot, *udata);
    if (check_flag(&retval, "external_forces (fEuler)", 1)) return -1;
  }
  retval = eulerb200_rhs_any(ctx, t, wp, wdp, NULL);
  if (retval != 0) {
    cerr << "\n" << eulerb200_last_error(ctx) << "\n\n";
    return -1;
  }
  // EULERB200_PROFILE=1: the reference's other profile slots of this path, fed from CUDA events
  // (device time of this call): PR_PACKDATA = per-cell pre-pass + halo pack kernels (the reference
  // times pack1D there, utilities.cpp:87-114), PR_MPI = halo transfer + the wait for it
  // (euler3D.hpp:600,789,1179), PR_FACEFLUX = interior kernel + boundary shells.
  static const bool profiling = getenv("EULERB200_PROFILE") != NULL && atoi(getenv("EULERB200_PROFILE")) != 0;
  if (profiling) {
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (eulerb200_profile(ctx, 1, 1, ms) == 0 && ms[7] > 0) {
      udata->profile[PR_PACKDATA].time += 1e-3 * (ms[1] + ms[2]);
      udata->profile[PR_PACKDATA].count++;
      udata->profile[PR_MPI].time += 1e-3 * (ms[3] + ms[5]);
      udata->profile[PR_MPI].count++;
      udata->profile[PR_FACEFLUX].time += 1e-3 * (ms[4] + ms[6]);
      udata->profile[PR_FACEFLUX].count++;
    }
  }

  retval = udata->profile[PR_RHSEULER].stop();
  if (check_flag(&retval, "Profile::stop (fEuler)", 1)) return -1;
  return 0;
}

int stability(N_Vector w, realtype t, realtype* dt_stab, void* user_data)
{
  (void)t;
  EulerData* udata = (EulerData*)user_data;
  int retval = udata->profile[PR_DTSTAB].start();
  if (check_flag(&retval, "Profile::start (stability)", 1)) return -1;
  double* wp[6];
  if (subvector_pointers(w, 0, "N_VGetSubvectorArrayPointer (stability)", wp)) return -1;
  eulerb200_ctx* ctx = context_for(udata);
  if (ctx == NULL) return -1;
  if (eulerb200_stability_any(ctx, wp, udata->cfl, dt_stab, NULL) != 0) {
    cerr << "\n" << eulerb200_last_error(ctx) << "\n\n";
    return -1;
  }
  retval = udata->profile[PR_DTSTAB].stop();
  if (check_flag(&retval, "Profile::stop (stability)", 1)) return -1;
  return 0;
}
