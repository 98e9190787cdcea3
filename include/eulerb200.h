/* ---------------------------------------------------------------------------
 * eulerb200.h -- C ABI of the B200-native fluid right-hand side
 *                f_1(w) = G - div F(w)   (5th-order FD-WENO, Lax-Friedrichs splitting)
 * of sundials-manyvector-demo.  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point names the reference interface it stands in for; paths are
 * relative to the reference tree, src/.
 *
 * Data layout (unchanged from the reference, euler3D.hpp:62,65):
 *   w[0..4] = rho, mx, my, mz, et   each nxl*nyl*nzl doubles, index i + nxl*(j + nyl*k)
 *   w[5]    = chem                  nxl*nyl*nzl*nchem doubles, index v + nchem*cell
 * i.e. exactly the six sub-vector arrays N_VGetSubvectorArrayPointer_MPIManyVector
 * returns (utilities.cpp:31-58).  Pointers are DEVICE (or managed) pointers unless the
 * function name ends in _host.
 *
 * Return codes follow the reference / ARKODE convention (utilities.cpp:542-591):
 *   0 success, <0 unrecoverable (-1: illegal state / invalid argument, -2: CUDA error,
 *   -3: communication error).  eulerb200_last_error() gives the text.
 * There is no CPU fallback: without a usable CUDA device eulerb200_create fails with -2.
 * ------------------------------------------------------------------------- */
#ifndef EULERB200_H
#define EULERB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EULERB200_VERSION 100

/* boundary condition codes: euler3D.hpp:90-93 */
#define EULERB200_BC_PERIODIC   0
#define EULERB200_BC_NEUMANN    1
#define EULERB200_BC_DIRICHLET  2
#define EULERB200_BC_REFLECTING 3

/* face order everywhere: W, E, S, N, B, F  (x-low, x-high, y-low, y-high, z-low, z-high) */
#define EULERB200_NO_NEIGHBOR (-1)   /* MPI_PROC_NULL: the face is a physical boundary */

typedef struct eulerb200_ctx eulerb200_ctx;

/* The fields of class EulerData the hot path reads (euler3D.hpp:177-270). */
typedef struct eulerb200_config {
  int64_t nxl, nyl, nzl;   /* local extents              euler3D.hpp:198-200 */
  int32_t nchem;           /* NVAR-5                      euler3D.hpp:220,297 */
  int32_t device;          /* CUDA device ordinal (-1: keep the current device) */
  double dx, dy, dz;       /* mesh spacing                euler3D.hpp:209-211 */
  double gamma;            /* ratio of specific heats     euler3D.hpp:221 */
  int32_t bc[6];           /* xlbc,xrbc,ylbc,yrbc,zlbc,zrbc   euler3D.hpp:214-219 */
  int32_t nbr[6];          /* ipW,ipE,ipS,ipN,ipB,ipF     euler3D.hpp:251-256;
                              EULERB200_NO_NEIGHBOR for a physical boundary; == rank when the
                              periodic wrap lands on this rank itself */
  int32_t rank, nranks;    /* myid, nprocs                euler3D.hpp:246-247 */
  double forcing[5];       /* constant G assigned into wdot: the external_forces hook
                              (euler3D.hpp:1454); zero for every shipped problem except
                              Rayleigh-Taylor's Gmy = -0.1 (rayleigh_taylor.cpp:117-128) */
} eulerb200_config;

int eulerb200_version(void);

/* EulerData::SetupDecomp, the arithmetic part (euler3D.hpp:416-440,466-567):
 * process grid from MPI_Dims_create over the axes with more than 3 cells, block extents
 * by integer division, neighbour ranks of a non-reordered Cartesian communicator.
 *   n[3] global cells, bc[6]; out: dims[3] = npx,npy,npz; coords[3];
 *   ext[6] = is,ie,js,je,ks,ke; nbr[6] = ipW..ipF (EULERB200_NO_NEIGHBOR = MPI_PROC_NULL).
 * Returns 0, 1 if only one boundary of an axis is periodic (euler3D.hpp:443-457),
 * -1 if a local extent would be < 3 (euler3D.hpp:483-494). */
int eulerb200_decompose(int32_t nprocs, int32_t rank, const int64_t* n, const int32_t* bc,
                        int32_t* dims, int32_t* coords, int64_t* ext, int32_t* nbr);

/* The point-to-point operations one halo exchange issues, in issue order: the role of the
 * Irecv/Isend tags of euler3D.hpp:608-640,663-784.  ops receives up to 12 triples
 * (kind, face, peer) with kind 0 = send, 1 = receive; returns their number. */
int eulerb200_exchange_plan(const eulerb200_config* cfg, int32_t* ops);

/* EulerData constructor + the allocation part of SetupDecomp (euler3D.hpp:497-559):
 * owns halo slabs, streams, events and the legal-state flag.  Never allocates afterwards. */
int eulerb200_create(const eulerb200_config* cfg, eulerb200_ctx** out);
/* EulerData::FreeData (euler3D.hpp:304-377) */
int eulerb200_destroy(eulerb200_ctx* ctx);
const char* eulerb200_last_error(const eulerb200_ctx* ctx);   /* ctx may be NULL: create errors */

/* Multi-GPU transport (stands in for the Cartesian communicator of euler3D.hpp:461 and
 * the Isend/Irecv pairs of euler3D.hpp:607-786).  One process per GPU.  Rank 0 obtains an
 * id, the host distributes the 128 bytes by whatever means it has (MPI_Bcast in the
 * reference drivers, torch.distributed in the tests), every rank attaches. */
#define EULERB200_UNIQUE_ID_BYTES 128
int eulerb200_comm_unique_id(void* id_bytes);
int eulerb200_comm_attach(eulerb200_ctx* ctx, const void* id_bytes);

/* Peer-store halo transport (one node, CUDA IPC over NVLink/NVSwitch), preferred over the
 * NCCL send/recv pairs when every neighbour is peer-accessible: the pack kernel of
 * ExchangeStart writes the three layers straight into the neighbour's ghost slab and
 * publishes a sequence number there; ExchangeEnd is a one-warp kernel that acquires it.
 * Every rank exports a blob, the host all-gathers the blobs in rank order (MPI_Allgather in
 * the drop-in, torch.distributed in the tests) and every rank attaches.  On failure (no
 * peer access) the context keeps using NCCL. */
#define EULERB200_P2P_BLOB_BYTES 256
int eulerb200_p2p_export(eulerb200_ctx* ctx, void* blob_bytes);
int eulerb200_p2p_attach(eulerb200_ctx* ctx, const void* all_blobs /* nranks blobs */);

/* fEuler (utilities.cpp:17-253): wdot = G - div F(w), including the halo exchange when
 * the context has neighbours.  Enqueued on `stream` (a cudaStream_t, NULL = default).
 * eulerb200_rhs returns only after the legal_state flag has been read back, like the
 * reference: 0, or -1 if any owned cell has rho<=0, et<=0 or p<=0 (utilities.cpp:83-84).
 * eulerb200_rhs_async does not synchronise; eulerb200_state_flag() later returns the
 * flag bits (1 density, 2 energy, 4 pressure; check_flag opt 4, utilities.cpp:575-588). */
int eulerb200_rhs(eulerb200_ctx* ctx, double t, const double* const* w, double* const* wdot,
                  void* stream);
int eulerb200_rhs_async(eulerb200_ctx* ctx, double t, const double* const* w, double* const* wdot,
                        void* stream);
int eulerb200_state_flag(eulerb200_ctx* ctx, void* stream, int32_t* bits);

/* The external_forces hook in full generality (euler3D.hpp:1454, called at utilities.cpp:65, where
 * it ASSIGNS G into wdot before the flux divergence is subtracted).  Constant-per-field forcing
 * lives in the config.  For a hook that varies in space or time, switch this on and run the hook
 * on wdot yourself before every eulerb200_rhs* call, exactly as the reference's fEuler does
 * (N_VConst(0, wdot); external_forces(t, wdot, udata)): the kernel then computes
 * wdot = wdot - div F(w) cell by cell, config.forcing is ignored, and eulerb200_rhs_host uploads
 * wdot along with w. */
int eulerb200_set_forcing_in_wdot(eulerb200_ctx* ctx, int32_t on);

/* fslow / fexpl of the multirate and IMEX drivers, fused (multirate_chem_hydro_main.cpp:
 * 996-1083, imex_chem_hydro_main.cpp:910-1000): rebuild the total energy from the gas energy
 * carried as the last chemistry species, et = chem[nchem-1]/EnergyUnits + |m|^2/(2 rho) -- written
 * into w[4] like the reference does -- evaluate fEuler, then chemdot[nchem-1] = etdot and
 * etdot = 0.  The Dengo scaling of chem before/after stays with the chemistry side.  Device
 * pointers; synchronous on the legal_state flag like eulerb200_rhs. */
int eulerb200_rhs_slow(eulerb200_ctx* ctx, double t, double* const* w, double* const* wdot,
                       double energy_units, void* stream);

/* Same call with HOST arrays (what a driver holding serial N_Vectors has): stages
 * host->device, evaluates, stages device->host, pipelined over z-slabs. */
int eulerb200_rhs_host(eulerb200_ctx* ctx, double t, const double* const* w_host,
                       double* const* wdot_host);

/* Dispatch on where w lives (device/managed -> eulerb200_rhs, host -> eulerb200_rhs_host);
 * what a drop-in fEuler that is handed arbitrary N_Vectors calls. */
int eulerb200_rhs_any(eulerb200_ctx* ctx, double t, const double* const* w, double* const* wdot,
                      void* stream);

/* EulerData::ExchangeStart / ExchangeEnd (euler3D.hpp:577-1191), callable on their own.
 * The packing runs on the library's exchange stream behind the work queued on `stream` so far:
 * w must stay unchanged until exchange_end has been queued on the stream that next writes it
 * (the reference packs inside ExchangeStart; fEuler, the only caller, never writes w). */
int eulerb200_exchange_start(eulerb200_ctx* ctx, const double* const* w, void* stream);
int eulerb200_exchange_end(eulerb200_ctx* ctx, void* stream);
/* Ghost layers of one face in the reference's receive-buffer layout (Wrecv..Frecv,
 * euler3D.hpp:257-262, index v + NVAR*(d + 3*(a + na*b))), whatever their source
 * (neighbour halo, periodic wrap, or boundary-condition fill euler3D.hpp:797-1166).
 * dst is a device pointer to eulerb200_face_len() doubles.  Call after exchange_end. */
int64_t eulerb200_face_len(const eulerb200_ctx* ctx, int32_t face);
int eulerb200_ghost_face(eulerb200_ctx* ctx, const double* const* w, int32_t face, double* dst,
                         void* stream);

/* stability (utilities.cpp:483-528): dt = cfl*min(dx,dy,dz)/max_cells(| |mx/rho| + c |),
 * max taken over all ranks.  Synchronous (returns the value). */
int eulerb200_stability(eulerb200_ctx* ctx, const double* const* w, double cfl, double* dt_stab,
                        void* stream);

/* Same, dispatching on where w lives (host arrays are staged to the device first). */
int eulerb200_stability_any(eulerb200_ctx* ctx, const double* const* w, double cfl, double* dt_stab,
                            void* stream);

/* Vector operations of the explicit driver loop (euler3D_main.cpp:357,387 hands these to
 * ARKODE, which evaluates them through N_VLinearCombination / N_VWrmsNorm on the
 * MPIManyVector); device pointers, one sub-vector per call.
 *   lincomb:     out[i] = sum_t coef[t] * x[t][i],  1 <= nterms <= 16, out may alias any x[t]
 *   wrms_accum:  *acc += sum_i (x[i] / (rtol*|y[i]| + atol))^2   (acc: device double) */
int eulerb200_vec_lincomb(eulerb200_ctx* ctx, int32_t nterms, const double* coef, const double* const* x,
                          double* out, int64_t n, void* stream);
int eulerb200_vec_wrms_accum(eulerb200_ctx* ctx, const double* x, const double* y, double rtol, double atol,
                             int64_t n, double* acc, void* stream);

/* N_VWrmsNorm over the whole ManyVector and all ranks (nglobal = global vector length);
 * synchronous, result on the host. */
int eulerb200_vec_wrms(eulerb200_ctx* ctx, const double* const* x, const double* const* y, double rtol,
                       double atol, int64_t nglobal, double* result, void* stream);

/* Device-memory helpers so that a C/C++ host driver needs no CUDA headers
 * (N_VNew_* / N_VDestroy and the host<->device copies of a device-vector build). */
void* eulerb200_device_alloc(int64_t bytes);
/* Managed memory (cudaMallocManaged): what the reference's own device builds use for the chemistry
 * sub-vector (N_VNewManaged_Raja, euler3D_main.cpp:158-166) -- host code of an unmodified driver
 * (initial conditions, diagnostics, I/O) reads and writes it, the RHS and the vector operations run
 * on it on the device.  Free with eulerb200_device_free. */
void* eulerb200_managed_alloc(int64_t bytes);
/* cudaDeviceSynchronize on the context's device: call before host code touches managed vectors. */
int eulerb200_synchronize(eulerb200_ctx* ctx);
void eulerb200_device_free(void* p);
int eulerb200_copy_to_device(void* dst, const void* src, int64_t bytes);
int eulerb200_copy_to_host(void* dst, const void* src, int64_t bytes);

/* Number of kernel launches issued through this context so far. */
int64_t eulerb200_launch_count(const eulerb200_ctx* ctx);

/* Device-time profile of the RHS calls since the last reset, from CUDA events on the streams the
 * phases run on -- the counterpart of the reference's Profile slots (euler3D.hpp:97-118:
 * PR_PACKDATA utilities.cpp:87-114, PR_MPI euler3D.hpp:600,789,1179, PR_FACEFLUX, PR_RHSEULER).
 * on != 0 switches the event recording on (it adds one stream synchronisation per RHS call);
 * out[8], averages in milliseconds per call: [0] whole call, [1] per-cell pre-pass (aux_kernel),
 * [2] halo pack kernels, [3] halo transfer (NCCL send/recv or peer stores, on the side stream),
 * [4] interior kernel, [5] wait for the halo after the interior kernel, [6] boundary shells,
 * [7] number of calls averaged.  out may be NULL (switch only). */
int eulerb200_profile(eulerb200_ctx* ctx, int32_t on, int32_t reset, double* out);

/* Measurement aid (no reference counterpart): sustained DFMA throughput of the current
 * device in TFLOP/s, the denominator of the FP64-pipe roofline bench.py reports. */
int eulerb200_fp64_peak(double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* EULERB200_H */
