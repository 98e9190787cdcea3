// ---------------------------------------------------------------------------
// eulerb200.cu -- implementation of the C ABI in include/eulerb200.h (sm_100a).
//
// Kernels in this translation unit:
//   rhs_fused_kernel     (rhs_kernel.cuh)  fEuler: faces + divergence, one pass
//   pack_face_kernel     (halo_kernels.cuh) halo pack of EulerData::ExchangeStart (euler3D.hpp:644-786)
//   ghost_face_kernel    (halo_kernels.cuh) materialise a face's ghost layers in the reference's
//                        receive-buffer layout (euler3D.hpp:797-1166; tests, drop-in)
//   wavespeed_kernel     (vector_kernels.cuh) local part of stability (utilities.cpp:505-513)
//   lincomb_kernel, wrms_kernel (vector_kernels.cuh) vector operations of the explicit driver loop
// There is deliberately no host implementation of any of them.
// ---------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define EB_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#include "host_setup.h"
#include "halo_kernels.cuh"
#include "vector_kernels.cuh"

namespace {

thread_local std::string g_create_error;

// ------------------------------- NCCL, resolved at run time -------------------------------
// (single-GPU use needs no NCCL at all; in a process that already loaded a libnccl.so.2,
// e.g. PyTorch's bundled one, dlopen returns that same library)
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl()
{
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
#define EB_SYM(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym); if (!api.field) return api
  EB_SYM(GetUniqueId, "ncclGetUniqueId");
  EB_SYM(CommInitRank, "ncclCommInitRank");
  EB_SYM(CommDestroy, "ncclCommDestroy");
  EB_SYM(Send, "ncclSend");
  EB_SYM(Recv, "ncclRecv");
  EB_SYM(AllReduce, "ncclAllReduce");
  EB_SYM(GroupStart, "ncclGroupStart");
  EB_SYM(GroupEnd, "ncclGroupEnd");
  EB_SYM(GetErrorString, "ncclGetErrorString");
#undef EB_SYM
  api.ok = true;
  return api;
}

// ----------------------------------- small kernels -----------------------------------

// ---- peer-store halo exchange (transport (b)): the pack kernel writes the
// three layers straight into the NEIGHBOUR's ghost slab over NVLink (CUDA IPC mapping of the
// neighbour's mailbox); a release store of the exchange's sequence number into the neighbour's
// arrival word follows in stream order, and the receiver's boundary kernels are preceded by a
// one-warp kernel that acquires it.  No host involvement, no staging buffer, no side stream.
__global__ void halo_signal_kernel(unsigned long long* peer_arrival, unsigned long long seq)
{
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_arrival), "l"(seq) : "memory");
}
__global__ void halo_wait_kernel(const unsigned long long* arrival, unsigned mask, unsigned long long seq,
                                 long long timeout_cycles, int* err)
{
  const int f = threadIdx.x;
  if (f < 6 && ((mask >> f) & 1u)) {
    const long long t0 = clock64();
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(arrival + f) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > timeout_cycles) { atomicOr(err, 8); break; }
      __nanosleep(200);
    } while (true);
  }
}

// DFMA throughput micro-benchmark: 8 independent dependency chains per thread.  Used by
// bench.py for the FP64-pipe roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 12345.678) out[0] = r;   // never true; keeps the chains alive
}

}  // namespace

// ------------------------------------- context -------------------------------------

struct eulerb200_ctx {
  eulerb200_config cfg;
  int device = 0;
  std::string error;
  int64_t launches = 0;

  int* d_flag = nullptr;
  int* h_flag = nullptr;                 // pinned
  unsigned long long* d_alpha = nullptr;
  double* h_alpha = nullptr;             // pinned

  bool remote[6] = {false, false, false, false, false, false};
  bool any_remote = false;
  double* send[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* recv[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // peer-store halo exchange (CUDA IPC): my mailbox = 2 parities x 6 ghost slabs + 6 arrival words
  bool p2p = false;
  char* mailbox = nullptr;
  int64_t slab_off[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};   // byte offsets inside the mailbox
  int64_t arrival_off = 0;
  char* peer_base[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t peer_slab_off[6][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};   // neighbour's slab for MY face f
  int64_t peer_arrival_off[6] = {0, 0, 0, 0, 0, 0};
  std::vector<void*> opened;
  unsigned long long seq = 0;
  const double* recv_cur[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_recv = nullptr;
  cudaStream_t slab_stream[3] = {nullptr, nullptr, nullptr};      // boundary slabs run side by side (highest stream priority)
  cudaEvent_t ev_fork = nullptr, ev_halo = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  bool exchange_open = false;

  // staging for eulerb200_rhs_host (allocated on first use)
  double* stage_w[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* stage_wdot[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t s_h2d = nullptr, s_cmp = nullptr, s_d2h = nullptr;
  static const int kMaxSlabs = 64;
  cudaEvent_t ev_up[kMaxSlabs], ev_done[kMaxSlabs];
  bool host_ready = false;
  size_t max_smem_set[2][6][3][3] = {}, carveout_for[2][6][3][3] = {};   // per kernel instantiation [xc][variant][kind][part]
  // eulerb200_profile: events T0 start, T1 after the pre-pass, T2 after the pack kernels, T3 after the interior
  // kernel, T4 after the wait for the halo, T5 end (stream s); C0 / C1 around the transfer (side stream)
  bool profile_on = false;
  cudaEvent_t pev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double pacc[7] = {0, 0, 0, 0, 0, 0, 0};
  double pacc_n = 0;
  bool forcing_in_wdot = false;  // eulerb200_set_forcing_in_wdot
  long ctas_target = 5920;       // EULERB200_CTAS: CTAs a launch aims for when cutting z-segments (tuning)
  int force_kernel = 1;          // EULERB200_KERNEL=0: never use the AG instantiation for boundary-heavy launches
  double ag_frac = 0.25;         // EULERB200_AG_FRAC: share of boundary tiles from which a launch takes the AG instantiation
  int variant = 0;
  int split = 0;                 // EULERB200_SPLIT=1: fluid fields and species in separate launches
  int stage = 0;                 // EULERB200_STAGE=1: the bulk-copy staging variant of the fused kernel (A/B)
  long long halo_timeout_cycles = 20000000000LL;   // peer-store halo wait: EULERB200_HALO_TIMEOUT_S (default 10 s) x the SM clock rate
  int overlap = 1;               // EULERB200_OVERLAP: 1 interior launch behind the halo exchange, then the shell launches;
                                 // 2 shells on high-priority streams as soon as the halo is in; 0 exchange first, one launch (rhs_impl)
  bool prof_pack_first = false;
  int thick_shells = 0;          // EULERB200_SHELLS=1: boundary shells one tile thick in x and y (default: three layers;
                                 // measured no faster: 60.42 vs 60.20 ms at 2 GPUs, the cost is in the boundary tiles themselves)
  int xc = 1;                    // EULERB200_XC=0: 31-column tiles (default: tiles own all 32 columns, the closing x-faces come
                                 // from the top warp; 59.2 vs 61.1 ms at 512^3 / NVAR 15)
  int variant_part[3] = {0, 0, 0};   // compiled variant per part (ALL, FLUID, TRACERS)
  double* aux[4] = {nullptr, nullptr, nullptr, nullptr};   // per-cell 1/rho, p, c, sqrt(rho)
  bool use_aux = true;
  double* chemT = nullptr;                                 // pair-interleaved copy of the species (RhsParams::chemT)
  bool use_chemT = false;                                  // EULERB200_CHEMT=1: species read from the pair-interleaved copy (measured:
                                                           // kernel -0.6 ms, pre-pass +5.4 ms at 512^3 / nchem 10, so off by default)
  int pair_sync = 2;                                       // EULERB200_PAIR: 0 CTA-wide barriers, 1/2 pairwise row rendezvous (rhs_kernel.cuh)
};

namespace {

// Compiled variants of the RHS kernel: threads per CTA and resident CTAs per SM fix the
// register budget (65536 / (threads * CTAs)).  EULERB200_VARIANT (fused launch) and
// EULERB200_VARIANT_F / _T (fluid / species launches of the split mode) select one by index for
// tuning runs; the defaults are the fastest measured on B200 (profiles/).
struct KernelVariant {
  // fn[kind][part].  kind 0: the default; 1: AG, boundary tiles read the per-cell arrays where valid
  // (for launches that are mostly boundary tiles); 2: GW, G taken from wdot (hook-assigned forcing).
  // part: eb::PART_ALL / PART_FLUID / PART_TRACERS.  nullptr: not compiled (launch_box falls back).
  void (*fn[3][3])(const eb::RhsParams);
  int threads;
  const char* name;
  // the XC instantiations of the same kinds and parts (rhs_fused_kernel<..., XC>), nullptr: not compiled
  void (*fnx[3][3])(const eb::RhsParams);
};
#define EB_K(T, B, GW, AG, PART, TYC) eb::rhs_fused_kernel<T, B, GW, AG, eb::PART, TYC>
#define EB_KX(T, B, GW, AG, PART, TYC) eb::rhs_fused_kernel<T, B, GW, AG, eb::PART, TYC, false, true>
#define EB_NONE {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}
// (XC instantiations: the fused launch of each kind; the split launches keep the 31-column tiles)
#ifdef EB_FAST_BUILD
#define EB_XC_ALL(T, B, TYC) {{EB_KX(T, B, false, false, PART_ALL, TYC), nullptr, nullptr}, {nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}
#else
#define EB_XC_ALL(T, B, TYC) {{EB_KX(T, B, false, false, PART_ALL, TYC), nullptr, nullptr}, {EB_KX(T, B, false, true, PART_ALL, TYC), nullptr, nullptr}, \
                              {EB_KX(T, B, true, false, PART_ALL, TYC), nullptr, nullptr}}
#endif
#define EB_KIND(T, B, GW, AG, TYC) {EB_K(T, B, GW, AG, PART_ALL, TYC), EB_K(T, B, GW, AG, PART_FLUID, TYC), EB_K(T, B, GW, AG, PART_TRACERS, TYC)}
#define EB_KERNELS_FULL(T, B, TYC) {EB_KIND(T, B, false, false, TYC), EB_KIND(T, B, false, true, TYC), EB_KIND(T, B, true, false, TYC)}
#define EB_KERNELS_PLAIN(T, B, TYC) {EB_KIND(T, B, false, false, TYC), {nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}
// [0..3]: CTAs of 32 x threads/32 threads, tile shape compiled in; [4]: any tile shape of up to 384
// threads (thin or small boxes, many species), shape read from blockDim
#ifdef EB_FAST_BUILD      // tuning builds: default kind only
#define EB_KERNELS_FULL EB_KERNELS_PLAIN
#endif
const KernelVariant kVariants[] = {
    {EB_KERNELS_PLAIN(256, 1, 8), 256, "256x1 (<=255 regs)", EB_NONE},
    {EB_KERNELS_FULL(384, 1, 12), 384, "384x1 (<=168 regs)", EB_XC_ALL(384, 1, 12)},
    {EB_KERNELS_FULL(512, 1, 16), 512, "512x1 (<=128 regs)", EB_NONE},      // (XC measured slower here: 28.4 vs 27.7 ms, RT 512^3)
    {EB_KERNELS_PLAIN(640, 1, 20), 640, "640x1 (<=96 regs)", EB_NONE},
    {EB_KERNELS_FULL(384, 1, 0), 384, "any tile shape, <= 384 threads", EB_NONE},
    // A/B variant (EULERB200_STAGE=1): species of the current plane staged in shared memory by cp.async.bulk
    {{{eb::rhs_fused_kernel<384, 1, false, false, eb::PART_ALL, 12, true>, nullptr, nullptr}, {nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}},
     384, "384x1, bulk-copy staging of the species plane", EB_NONE},
};
const int kGenericVariant = 4;
const int kStageVariant = 5;
const int kNumVariants = 4;    // selectable; the last entry is the any-shape fallback
const int kDefaultVariant = 1;

int fail(eulerb200_ctx* c, int code, const std::string& msg)
{
  if (c) c->error = msg; else g_create_error = msg;
  return code;
}
#define EB_CUDA(c, call)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(c, -2, std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #call); \
  } while (0)
#define EB_NCCL(c, call)                                                                  \
  do {                                                                                    \
    ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess)                                                                \
      return fail(c, -3, std::string("NCCL error: ") + nccl().GetErrorString(r_) + " in " #call); \
  } while (0)

// eulerb200_profile: record event q on stream st when profiling is on
#define EB_PREC(c, q, st) do { if ((c)->profile_on) EB_CUDA(c, cudaEventRecord((c)->pev[q], st)); } while (0)

eb::RhsParams make_params(eulerb200_ctx* c, const double* const* w, double* const* wdot)
{
  const eulerb200_config& g = c->cfg;
  eb::RhsParams P;
  P.nx = g.nxl; P.ny = g.nyl; P.nz = g.nzl;
  P.nchem = g.nchem;
  P.gamma = g.gamma;
  P.rdx = EB_RD_SCALE / g.dx; P.rdy = EB_RD_SCALE / g.dy; P.rdz = EB_RD_SCALE / g.dz;
  P.dx = g.dx; P.dy = g.dy; P.dz = g.dz;
  for (int f = 0; f < 5; f++) P.forcing[f] = g.forcing[f];
  for (int f = 0; f < 6; f++) {
    P.w[f] = (f < 5 || g.nchem > 0) ? w[f] : nullptr;
    P.wdot[f] = (f < 5 || g.nchem > 0) ? wdot[f] : nullptr;
    eb::ghost_face(g, f, c->recv_cur[f], &P.ghost[f]);
  }
  for (int q = 0; q < 4; q++) P.aux[q] = c->use_aux ? c->aux[q] : nullptr;
  P.chemT = (c->use_chemT && g.nchem > 0) ? c->chemT : nullptr;
  P.slow_mode = 0;
  P.vec_store = (g.nchem > 0 && (g.nchem & 1) == 0 && (((unsigned long long)wdot[5]) & 15ull) == 0ull) ? 1 : 0;
  P.pair_sync = 0;
  P.inv_energy_units = 1.0;
  P.et_rw = nullptr;
  P.state_flag = c->d_flag;
  P.lo[0] = P.lo[1] = P.lo[2] = 0;
  P.hi[0] = P.nx; P.hi[1] = P.ny; P.hi[2] = P.nz;
  P.seg_len = 1;
  return P;
}

// Evaluate the cells of [lo,hi) (clipped to non-empty) on `s`: one fused launch, or (split mode,
// nchem > 0) a fluid launch followed by a species launch.
int launch_part(eulerb200_ctx* c, eb::RhsParams P, int kind, int part, cudaStream_t s)
{
  const int nf = part == eb::PART_ALL ? 5 + P.nchem : (part == eb::PART_FLUID ? 5 : P.nchem);
  int vi = c->variant_part[part];
  eb::LaunchGeom L = eb::launch_geom(P.lo, P.hi, nf, kVariants[vi].threads, c->pair_sync, c->ctas_target);
  if (L.tx != 32 || L.ty != kVariants[vi].threads / 32 || !kVariants[vi].fn[kind][part]) {
    // not the compiled-in tile shape (thin or small box, many species), or a kind compiled for the
    // default variant only: the any-shape instantiation
    vi = kGenericVariant;
    L = eb::launch_geom(P.lo, P.hi, nf, kVariants[vi].threads, c->pair_sync, c->ctas_target);
  }
  if (c->stage && kind == 0 && part == eb::PART_ALL && vi == kDefaultVariant && P.vec_store && !P.chemT) {
    // the staging variant: CTA-wide barriers, and room for the (TX+5) x (TY+5) x nchem window + the mbarrier
    const size_t extra = sizeof(double) * (size_t)(L.tx + 5) * (L.ty + 5) * P.nchem + 16;
    eb::LaunchGeom Ls = eb::launch_geom(P.lo, P.hi, nf, kVariants[kStageVariant].threads, 0, c->ctas_target);
    if (Ls.tx == L.tx && Ls.ty == L.ty && Ls.smem + extra <= (size_t)227 * 1024) {
      vi = kStageVariant;
      L = Ls;
      L.smem += extra;
    }
  }
  const KernelVariant& V = kVariants[vi];
  void (*fn)(const eb::RhsParams) = V.fn[kind][part];
  if (c->xc && V.fnx[kind][part] && vi != kStageVariant) {
    // tiles of 32 owned columns (the launch keeps the compiled-in 32 x threads/32 shape: checked above)
    const eb::LaunchGeom Lx = eb::launch_geom(P.lo, P.hi, nf, V.threads, c->pair_sync, c->ctas_target, 1);
    if (Lx.xc && Lx.tx == L.tx && Lx.ty == L.ty && Lx.smem <= (size_t)227 * 1024) {
      L = Lx;
      fn = V.fnx[kind][part];
    }
  }
  P.seg_len = L.seg_len;
  P.pair_sync = L.pair;
  size_t& smem_set = c->max_smem_set[L.xc][vi][kind][part];
  size_t& carve_for = c->carveout_for[L.xc][vi][kind][part];
  if (L.smem > smem_set) {
    EB_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    smem_set = L.smem;
  }
  if (L.smem != carve_for) {
    // the stencil loads live in L1: ask for the smallest shared-memory carve-out that holds one CTA
    // (+1 KB the system reserves per CTA) and leave the rest of the 256 KB to L1
    // (the carve-outs sm_100 offers; asked for as floor(100 S / 228), the form the 132 KB one was measured with)
    static const size_t kCarve[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
    size_t pick = 228;
    for (size_t q = 0; q < sizeof kCarve / sizeof kCarve[0]; q++)
      if (kCarve[q] * 1024 >= L.smem + 1024) { pick = kCarve[q]; break; }
    const int pct = (int)(100 * pick / 228);
    EB_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    carve_for = L.smem;
  }
  fn<<<dim3(L.gx, L.gy, L.gz), dim3(L.tx, L.ty, 1), L.smem, s>>>(P);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  return 0;
}

int launch_box(eulerb200_ctx* c, eb::RhsParams P, const long lo[3], const long hi[3], cudaStream_t s)
{
  for (int d = 0; d < 3; d++) {
    if (hi[d] <= lo[d]) return 0;
    P.lo[d] = lo[d]; P.hi[d] = hi[d];
  }
  // which instantiation: hook-assigned forcing -> kind 2; with EULERB200_KERNEL=1 a launch in which a
  // quarter or more of the tiles touch a boundary (thin or small grids, the boundary shells of a
  // decomposed run) -> kind 1; else the default kind 0
  int kind = 0;
  if (c->forcing_in_wdot) kind = 2;
  else if (c->force_kernel == 1 && c->use_aux) {
    const eb::LaunchGeom L = eb::launch_geom(P.lo, P.hi, 5 + P.nchem, kVariants[c->variant_part[0]].threads, c->pair_sync, c->ctas_target, c->xc);
    if (eb::boundary_tile_fraction(P.lo, P.hi, P.nx, P.ny, L) >= c->ag_frac) kind = 1;
  }
  if (!kVariants[kGenericVariant].fn[kind][0]) {      // reduced builds (tuning, strict) compile the default kind only
    if (kind == 2) return fail(c, -1, "this build of the library has no hook-assigned-forcing instantiation");
    kind = 0;
  }
  int rc;
  if (c->split && P.nchem > 0) {
    rc = launch_part(c, P, kind, eb::PART_FLUID, s);
    if (!rc) rc = launch_part(c, P, kind, eb::PART_TRACERS, s);
  } else {
    rc = launch_part(c, P, kind, eb::PART_ALL, s);
  }
  if (rc || !P.slow_mode) return rc;
  // fslow / fexpl: etdot moves into the gas-energy species of this box
  const long ncell = (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  eb::slow_post_kernel<<<(unsigned)std::min<long>((ncell + 255) / 256, 148L * 16), 256, 0, s>>>(P);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  return 0;
}

// Per-cell derived values for the z-planes [k0, k1) of the state in P.
int launch_aux(eulerb200_ctx* c, const eb::RhsParams& P, long k0, long k1, cudaStream_t s)
{
  if ((!c->use_aux && !P.slow_mode && !P.chemT) || k1 <= k0) return 0;
  const long plane = P.nx * P.ny, c0 = k0 * plane, c1 = k1 * plane;
  const unsigned blocks = (unsigned)std::min<long>((c1 - c0 + 255) / 256, 148L * 16);
  eb::aux_kernel<<<blocks, 256, 0, s>>>(P, c->aux[0], c->aux[1], c->aux[2], c->aux[3], const_cast<double*>(P.chemT), c0, c1);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  return 0;
}

eb::FaceGeom face_geom(const eulerb200_ctx* c, int f, const double* const* w)
{
  eb::FaceGeom g;
  g.nx = c->cfg.nxl; g.ny = c->cfg.nyl; g.nz = c->cfg.nzl;
  g.nchem = c->cfg.nchem; g.f = f;
  for (int q = 0; q < 6; q++) g.w[q] = (q < 5 || c->cfg.nchem > 0) ? w[q] : nullptr;
  return g;
}

int exchange_start(eulerb200_ctx* c, const double* const* w, cudaStream_t s)
{
  if (!c->any_remote) return 0;
  const int nv = 5 + c->cfg.nchem;
  if (c->p2p) {
    // pack straight into the neighbours' ghost slabs (parity = exchange number mod 2: the slab a
    // neighbour may still be reading belongs to the previous exchange), then publish the number
    c->seq++;
    const int par = (int)(c->seq & 1ull);
    // on the side stream, so that the NVLink stores overlap the interior kernel instead of
    // delaying it (the pack CTAs slip into the SMs as interior CTAs retire)
    EB_CUDA(c, cudaEventRecord(c->ev_packed, s));
    EB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_packed, 0));
    EB_PREC(c, 6, c->comm_stream);
    for (int f = 0; f < 6; f++) {
      if (!c->remote[f]) continue;
      const long nent = eb::face_len(c->cfg, f) / nv;
      double* dst = reinterpret_cast<double*>(c->peer_base[f] + c->peer_slab_off[f][par]);
      eb::pack_face_warp_kernel<<<(unsigned)((nent + 7) / 8), dim3(32, 8), 0, c->comm_stream>>>(face_geom(c, f, w), dst, nent);
      halo_signal_kernel<<<1, 1, 0, c->comm_stream>>>(
          reinterpret_cast<unsigned long long*>(c->peer_base[f] + c->peer_arrival_off[f]), c->seq);
      c->launches += 2;
      c->recv_cur[f] = reinterpret_cast<const double*>(c->mailbox + c->slab_off[par][f]);
    }
    EB_CUDA(c, cudaGetLastError());
    EB_PREC(c, 7, c->comm_stream);
    EB_CUDA(c, cudaEventRecord(c->ev_recv, c->comm_stream));
    c->exchange_open = true;
    return 0;
  }
  if (!c->comm) return fail(c, -3, "context has remote neighbours but neither eulerb200_comm_attach nor eulerb200_p2p_attach was called");
  // pack on the exchange stream as well: it only needs w as it is when this call is made (everything queued
  // on s so far), so nothing the caller launches on s afterwards -- the pre-pass, the interior -- waits for it
  EB_CUDA(c, cudaEventRecord(c->ev_packed, s));
  EB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_packed, 0));
  EB_PREC(c, 6, c->comm_stream);
  for (int f = 0; f < 6; f++) {
    if (!c->remote[f]) continue;
    const long nent = eb::face_len(c->cfg, f) / nv;
    eb::pack_face_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, c->comm_stream>>>(face_geom(c, f, w), c->send[f], nent);
    c->launches++;
  }
  EB_CUDA(c, cudaGetLastError());
  eb::ExchangeOp ops[12];
  const int nops = eb::exchange_plan(c->cfg, ops);
  EB_NCCL(c, nccl().GroupStart());
  for (int q = 0; q < nops; q++) {
    const int f = ops[q].face;
    const size_t len = (size_t)eb::face_len(c->cfg, f);
    if (ops[q].kind == 0) EB_NCCL(c, nccl().Send(c->send[f], len, ncclDouble, ops[q].peer, c->comm, c->comm_stream));
    else EB_NCCL(c, nccl().Recv(c->recv[f], len, ncclDouble, ops[q].peer, c->comm, c->comm_stream));
  }
  EB_NCCL(c, nccl().GroupEnd());
  EB_PREC(c, 7, c->comm_stream);
  EB_CUDA(c, cudaEventRecord(c->ev_recv, c->comm_stream));
  c->exchange_open = true;
  return 0;
}

int exchange_end(eulerb200_ctx* c, cudaStream_t s)
{
  if (!c->exchange_open) return 0;
  if (c->p2p) {
    EB_CUDA(c, cudaStreamWaitEvent(s, c->ev_recv, 0));     // my own stores are out (w may be reused)
    unsigned mask = 0;
    for (int f = 0; f < 6; f++) if (c->remote[f]) mask |= 1u << f;
    halo_wait_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(c->mailbox + c->arrival_off), mask,
                                      c->seq, c->halo_timeout_cycles, c->d_flag);
    c->launches++;
    EB_CUDA(c, cudaGetLastError());
  } else {
    EB_CUDA(c, cudaStreamWaitEvent(s, c->ev_recv, 0));
  }
  c->exchange_open = false;
  return 0;
}

int ensure_staging(eulerb200_ctx* c)
{
  if (c->host_ready) return 0;
  const eulerb200_config& g = c->cfg;
  const long N = g.nxl * g.nyl * g.nzl;
  const int nsub = 5 + (g.nchem > 0 ? 1 : 0);
  for (int f = 0; f < nsub; f++) {
    const size_t bytes = sizeof(double) * N * (f < 5 ? 1 : g.nchem);
    EB_CUDA(c, cudaMalloc(&c->stage_w[f], bytes));
    EB_CUDA(c, cudaMalloc(&c->stage_wdot[f], bytes));
  }
  EB_CUDA(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  EB_CUDA(c, cudaStreamCreateWithFlags(&c->s_cmp, cudaStreamNonBlocking));
  EB_CUDA(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  for (int s = 0; s < eulerb200_ctx::kMaxSlabs; s++) {
    EB_CUDA(c, cudaEventCreateWithFlags(&c->ev_up[s], cudaEventDisableTiming));
    EB_CUDA(c, cudaEventCreateWithFlags(&c->ev_done[s], cudaEventDisableTiming));
  }
  c->host_ready = true;
  return 0;
}

}  // namespace

// ------------------------------------- C ABI -------------------------------------

extern "C" {

int eulerb200_version(void) { return EULERB200_VERSION; }

int eulerb200_decompose(int32_t nprocs, int32_t rank, const int64_t* n, const int32_t* bc,
                        int32_t* dims, int32_t* coords, int64_t* ext, int32_t* nbr)
{
  return eb::decompose(nprocs, rank, n, bc, dims, coords, ext, nbr);
}

int eulerb200_exchange_plan(const eulerb200_config* cfg, int32_t* ops)
{
  if (!cfg || !ops) return -1;
  eb::ExchangeOp tmp[12];
  const int n = eb::exchange_plan(*cfg, tmp);
  for (int q = 0; q < n; q++) { ops[3 * q] = tmp[q].kind; ops[3 * q + 1] = tmp[q].face; ops[3 * q + 2] = tmp[q].peer; }
  return n;
}

const char* eulerb200_last_error(const eulerb200_ctx* ctx)
{
  return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

int eulerb200_create(const eulerb200_config* cfg, eulerb200_ctx** out)
{
  if (!cfg || !out) return fail(nullptr, -1, "null argument");
  *out = nullptr;
  if (cfg->nxl < 3 || cfg->nyl < 3 || cfg->nzl < 3) return fail(nullptr, -1, "local extents must be >= 3 (euler3D.hpp:483-494)");
  if (cfg->nchem < 0 || 5 + cfg->nchem > 64) return fail(nullptr, -1, "nchem out of range");
  if ((double)cfg->nxl * (double)cfg->nyl * (double)cfg->nzl >= 2147483648.0) return fail(nullptr, -1, "more than 2^31 - 1 cells per rank");
  if (!(cfg->dx > 0) || !(cfg->dy > 0) || !(cfg->dz > 0)) return fail(nullptr, -1, "mesh spacing must be positive");
  for (int f = 0; f < 6; f++) {
    eb::GhostFace g;
    if (eb::ghost_face(*cfg, f, nullptr, &g) != 0)
      return fail(nullptr, -1, "periodic face without a neighbour (set nbr[f] = rank for a wrap onto this rank)");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, -2, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  eulerb200_ctx* c = new eulerb200_ctx();
  c->cfg = *cfg;
  // 384 threads (168 registers) with species; without them the flux arrays are small, the 128-register
  // build spills little and 16 warps per SM win (measured at 512^3: 27.6 vs 29.3 ms; with ten species
  // 65.4 vs 61.7 ms the other way round, profiles/README.md)
  c->variant = (cfg->nchem == 0) ? 2 : kDefaultVariant;
  if (const char* ev = getenv("EULERB200_VARIANT")) {
    const int v = atoi(ev);
    if (v >= 0 && v < kNumVariants) c->variant = v;
  }
  c->variant_part[0] = c->variant;
  c->variant_part[1] = c->variant_part[2] = kDefaultVariant;
  if (const char* ev = getenv("EULERB200_VARIANT_F")) { const int v = atoi(ev); if (v >= 0 && v < kNumVariants) c->variant_part[1] = v; }
  if (const char* ev = getenv("EULERB200_VARIANT_T")) { const int v = atoi(ev); if (v >= 0 && v < kNumVariants) c->variant_part[2] = v; }
  if (const char* ev = getenv("EULERB200_SPLIT")) c->split = atoi(ev) != 0;
  if (const char* ev = getenv("EULERB200_STAGE")) c->stage = atoi(ev) != 0;
  if (const char* ev = getenv("EULERB200_XC")) c->xc = atoi(ev) != 0;
  if (const char* ev = getenv("EULERB200_SHELLS")) c->thick_shells = atoi(ev) != 0;
  if (const char* ev = getenv("EULERB200_OVERLAP")) c->overlap = std::max(0, std::min(2, atoi(ev)));
  for (int x_ = 0; x_ < 2; x_++) for (int a_ = 0; a_ < 6; a_++) for (int b_ = 0; b_ < 3; b_++) for (int d_ = 0; d_ < 3; d_++) c->carveout_for[x_][a_][b_][d_] = (size_t)-1;
  if (cfg->device >= 0) {
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { delete c; return fail(nullptr, -2, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
  }
  cudaGetDevice(&c->device);
  {
    // how long a rank may wait for a neighbour's ghost layers before the right-hand side fails with -3 (a rank
    // stalled for longer, e.g. on I/O, is a hard failure): seconds from the environment, cycles from the clock rate
    double secs = 10.0;
    if (const char* ev = getenv("EULERB200_HALO_TIMEOUT_S")) secs = std::max(0.001, atof(ev));
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device) != cudaSuccess || khz <= 0) khz = 2000000;
    c->halo_timeout_cycles = (long long)(secs * 1e3 * (double)khz);
  }
#define EB_CREATE(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      std::string m_ = std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #call; \
      eulerb200_destroy(c);                                                               \
      return fail(nullptr, -2, m_);                                                       \
    }                                                                                     \
  } while (0)
  EB_CREATE(cudaMalloc(&c->d_flag, sizeof(int)));
  EB_CREATE(cudaMemset(c->d_flag, 0, sizeof(int)));
  EB_CREATE(cudaMallocHost(&c->h_flag, sizeof(int)));
  EB_CREATE(cudaMalloc(&c->d_alpha, sizeof(unsigned long long)));
  EB_CREATE(cudaMallocHost(&c->h_alpha, sizeof(double)));
  if (const char* ev = getenv("EULERB200_NO_AUX")) c->use_aux = (atoi(ev) == 0);
  if (const char* ev = getenv("EULERB200_PAIR")) c->pair_sync = std::max(0, std::min(2, atoi(ev)));
  if (const char* ev = getenv("EULERB200_KERNEL")) c->force_kernel = (atoi(ev) != 0) ? 1 : 0;
  if (const char* ev = getenv("EULERB200_AG_FRAC")) c->ag_frac = atof(ev);
  if (const char* ev = getenv("EULERB200_CTAS")) c->ctas_target = std::max(1L, atol(ev));
  if (const char* ev = getenv("EULERB200_CHEMT")) c->use_chemT = (atoi(ev) != 0);
  if (c->use_aux)
    for (int q = 0; q < 4; q++)
      EB_CREATE(cudaMalloc(&c->aux[q], sizeof(double) * cfg->nxl * cfg->nyl * cfg->nzl));
  if (c->use_chemT && cfg->nchem > 0)
    EB_CREATE(cudaMalloc(&c->chemT, sizeof(double) * 2 * ((cfg->nchem + 1) / 2) * cfg->nxl * cfg->nyl * cfg->nzl));
  for (int f = 0; f < 6; f++) {
    c->remote[f] = eb::face_is_remote(*cfg, f);
    if (c->remote[f]) {
      c->any_remote = true;
      EB_CREATE(cudaMalloc(&c->send[f], sizeof(double) * eb::face_len(*cfg, f)));
      EB_CREATE(cudaMalloc(&c->recv[f], sizeof(double) * eb::face_len(*cfg, f)));
      c->recv_cur[f] = c->recv[f];
    }
  }
  if (c->any_remote) {
    // The exchange stream and the shell streams get the highest priority: their kernels are small and the
    // interior launch keeps every SM occupied (one 384-thread CTA holds the whole register file), so at
    // equal priority the pack / NCCL copy CTAs are only scheduled when the interior grid runs dry -- the
    // halo then arrives after 57 ms instead of 2 (timeline of profiles/r2_bench_n2.json).
    int prio_least = 0, prio_greatest = 0;
    EB_CREATE(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    EB_CREATE(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_greatest));
    EB_CREATE(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    EB_CREATE(cudaEventCreateWithFlags(&c->ev_recv, cudaEventDisableTiming));
    EB_CREATE(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    EB_CREATE(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    for (int h = 0; h < 3; h++) {
      EB_CREATE(cudaStreamCreateWithPriority(&c->slab_stream[h], cudaStreamNonBlocking, prio_greatest));
      EB_CREATE(cudaEventCreateWithFlags(&c->ev_join[h], cudaEventDisableTiming));
    }
  }
#undef EB_CREATE
  *out = c;
  return 0;
}

int eulerb200_destroy(eulerb200_ctx* c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->opened) cudaIpcCloseMemHandle(p);
  if (c->mailbox) cudaFree(c->mailbox);
  if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
  for (int f = 0; f < 6; f++) {
    if (c->send[f]) cudaFree(c->send[f]);
    if (c->recv[f]) cudaFree(c->recv[f]);
    if (c->stage_w[f]) cudaFree(c->stage_w[f]);
    if (c->stage_wdot[f]) cudaFree(c->stage_wdot[f]);
  }
  if (c->host_ready) {
    for (int s = 0; s < eulerb200_ctx::kMaxSlabs; s++) { cudaEventDestroy(c->ev_up[s]); cudaEventDestroy(c->ev_done[s]); }
    cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_cmp); cudaStreamDestroy(c->s_d2h);
  }
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->ev_packed) cudaEventDestroy(c->ev_packed);
  if (c->ev_recv) cudaEventDestroy(c->ev_recv);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  for (int h = 0; h < 3; h++) {
    if (c->slab_stream[h]) cudaStreamDestroy(c->slab_stream[h]);
    if (c->ev_join[h]) cudaEventDestroy(c->ev_join[h]);
  }
  for (int q = 0; q < 4; q++) if (c->aux[q]) cudaFree(c->aux[q]);
  for (int q = 0; q < 8; q++) if (c->pev[q]) cudaEventDestroy(c->pev[q]);
  if (c->chemT) cudaFree(c->chemT);
  if (c->d_flag) cudaFree(c->d_flag);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  if (c->d_alpha) cudaFree(c->d_alpha);
  if (c->h_alpha) cudaFreeHost(c->h_alpha);
  delete c;
  return 0;
}

int eulerb200_comm_unique_id(void* id_bytes)
{
  if (!nccl().ok) return fail(nullptr, -3, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, -3, std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == EULERB200_UNIQUE_ID_BYTES, "unique id size");
  memcpy(id_bytes, &id, sizeof id);
  return 0;
}

int eulerb200_comm_attach(eulerb200_ctx* c, const void* id_bytes)
{
  if (!c) return -1;
  if (!nccl().ok) return fail(c, -3, "libnccl.so.2 could not be loaded");
  EB_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  EB_NCCL(c, nccl().CommInitRank(&c->comm, c->cfg.nranks, id, c->cfg.rank));
  return 0;
}

// Blob a rank publishes: its IPC handle, where each ghost slab sits in its mailbox, and where
// the arrival words are.
struct P2PBlob {
  cudaIpcMemHandle_t handle;                 // 64 bytes
  int64_t slab_off[2][6];
  int64_t arrival_off;
  int64_t face_len[6];
  int32_t rank, valid;
  char pad[EULERB200_P2P_BLOB_BYTES - 64 - 96 - 8 - 48 - 8];
};
static_assert(sizeof(P2PBlob) == EULERB200_P2P_BLOB_BYTES, "blob size");

int eulerb200_p2p_export(eulerb200_ctx* c, void* blob_bytes)
{
  if (!c || !blob_bytes) return -1;
  EB_CUDA(c, cudaSetDevice(c->device));
  P2PBlob b;
  memset(&b, 0, sizeof b);
  b.rank = c->cfg.rank;
  if (!c->mailbox) {
    int64_t off = 0;
    for (int par = 0; par < 2; par++)
      for (int f = 0; f < 6; f++) {
        c->slab_off[par][f] = off;
        if (c->remote[f]) off += ((int64_t)sizeof(double) * eb::face_len(c->cfg, f) + 255) / 256 * 256;
      }
    c->arrival_off = off;
    off += 256;
    EB_CUDA(c, cudaMalloc(&c->mailbox, (size_t)off));
    EB_CUDA(c, cudaMemset(c->mailbox, 0, (size_t)off));
  }
  EB_CUDA(c, cudaIpcGetMemHandle(&b.handle, c->mailbox));
  memcpy(b.slab_off, c->slab_off, sizeof b.slab_off);
  b.arrival_off = c->arrival_off;
  for (int f = 0; f < 6; f++) b.face_len[f] = c->remote[f] ? eb::face_len(c->cfg, f) : 0;
  b.valid = 1;
  memcpy(blob_bytes, &b, sizeof b);
  return 0;
}

int eulerb200_p2p_attach(eulerb200_ctx* c, const void* all_blobs)
{
  if (!c || !all_blobs) return -1;
  if (!c->mailbox) return fail(c, -3, "eulerb200_p2p_export must be called first");
  EB_CUDA(c, cudaSetDevice(c->device));
  const P2PBlob* blobs = reinterpret_cast<const P2PBlob*>(all_blobs);
  std::vector<char*> base(c->cfg.nranks, nullptr);
  for (int f = 0; f < 6; f++) {
    if (!c->remote[f]) continue;
    const int r = c->cfg.nbr[f];
    const P2PBlob& b = blobs[r];
    if (!b.valid || b.rank != r) return fail(c, -3, "peer blob missing or out of order");
    if (b.face_len[f ^ 1] != eb::face_len(c->cfg, f)) return fail(c, -3, "neighbour's ghost slab has a different size");
    if (!base[r]) {
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(c, -3, std::string("cudaIpcOpenMemHandle failed (no peer access?): ") + cudaGetErrorString(e));
      }
      base[r] = (char*)p;
      c->opened.push_back(p);
    }
    c->peer_base[f] = base[r];
    // what I send through my face f lands in the neighbour's slab of the opposite face
    for (int par = 0; par < 2; par++) c->peer_slab_off[f][par] = b.slab_off[par][f ^ 1];
    c->peer_arrival_off[f] = b.arrival_off + (int64_t)sizeof(unsigned long long) * (f ^ 1);
  }
  c->p2p = true;
  return 0;
}

int eulerb200_exchange_start(eulerb200_ctx* c, const double* const* w, void* stream)
{
  if (!c || !w) return -1;
  return exchange_start(c, w, (cudaStream_t)stream);
}

int eulerb200_exchange_end(eulerb200_ctx* c, void* stream)
{
  if (!c) return -1;
  return exchange_end(c, (cudaStream_t)stream);
}

int64_t eulerb200_face_len(const eulerb200_ctx* c, int32_t face)
{
  return (c && face >= 0 && face < 6) ? eb::face_len(c->cfg, face) : -1;
}

int eulerb200_ghost_face(eulerb200_ctx* c, const double* const* w, int32_t f, double* dst, void* stream)
{
  if (!c || !w || !dst || f < 0 || f >= 6) return -1;
  eb::GhostFace G;
  eb::ghost_face(c->cfg, f, c->recv_cur[f], &G);
  const long nent = eb::face_len(c->cfg, f) / (5 + c->cfg.nchem);
  eb::ghost_face_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, (cudaStream_t)stream>>>(face_geom(c, f, w), G, dst, nent);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  return 0;
}

static int rhs_impl(eulerb200_ctx* c, const double* const* w, double* const* wdot, void* stream,
                    int slow_mode, double energy_units);

int eulerb200_rhs_async(eulerb200_ctx* c, double t, const double* const* w, double* const* wdot, void* stream)
{
  (void)t;   // every shipped forcing is time independent
  return rhs_impl(c, w, wdot, stream, 0, 1.0);
}

// What the device flag of a right-hand side says, for every entry point that reads it back (rhs, rhs_slow,
// rhs_host; the async path hands the raw bits to eulerb200_state_flag's caller): bit 8 = a rank's peer-store
// halo wait timed out (communication error, -3), bits 1 / 2 / 4 = legal_state (euler3D.hpp:1405-1414, -1).
static int decode_flag(eulerb200_ctx* c, int32_t bits)
{
  if (bits & 8) return fail(c, -3, "halo exchange timed out waiting for a neighbour's ghost layers (peer-store transport)");
  if (bits) {
    char msg[160];
    snprintf(msg, sizeof msg, "STATE_ERROR: legal_state (fEuler) failed with flag = %d (1 density, 2 energy, 4 pressure)", bits);
    return fail(c, -1, msg);
  }
  return 0;
}

int eulerb200_rhs_slow(eulerb200_ctx* c, double t, double* const* w, double* const* wdot, double energy_units,
                       void* stream)
{
  (void)t;
  if (!c) return -1;
  if (c->cfg.nchem < 1) return fail(c, -1, "slow RHS needs the gas energy as the last chemistry species (nchem >= 1)");
  if (!(energy_units > 0)) return fail(c, -1, "EnergyUnits must be positive (euler3D.hpp:385-393)");
  int rc = rhs_impl(c, w, wdot, stream, 1, energy_units);
  if (rc) return rc;
  int32_t bits = 0;
  rc = eulerb200_state_flag(c, stream, &bits);
  if (rc) return rc;
  return decode_flag(c, bits);
}

static int profile_collect(eulerb200_ctx* c)
{
  if (!c->profile_on) return 0;
  EB_CUDA(c, cudaEventSynchronize(c->pev[5]));
  auto el = [&](int a, int b) { float ms = 0; if (cudaEventElapsedTime(&ms, c->pev[a], c->pev[b]) != cudaSuccess) { cudaGetLastError(); ms = 0; } return (double)ms; };
  c->pacc[0] += el(0, 5);
  c->pacc[c->prof_pack_first ? 2 : 1] += el(0, 1);
  if (c->any_remote) {
    c->pacc[c->prof_pack_first ? 1 : 2] += el(1, 2);
    if (cudaEventSynchronize(c->pev[7]) == cudaSuccess) c->pacc[3] += el(6, 7);
    c->pacc[4] += el(2, 3);
    c->pacc[5] += el(3, 4);
    c->pacc[6] += el(4, 5);
  } else {
    c->pacc[4] += el(1, 5);
  }
  c->pacc_n += 1.0;
  return 0;
}

static int rhs_impl(eulerb200_ctx* c, const double* const* w, double* const* wdot, void* stream,
                    int slow_mode, double energy_units)
{
  if (!c || !w || !wdot) return -1;
  EB_CUDA(c, cudaSetDevice(c->device));        // one context per GPU; the caller may have moved on
  for (int f = 0; f < 5 + (c->cfg.nchem > 0 ? 1 : 0); f++)
    if (!w[f] || !wdot[f]) return fail(c, -1, "NULL sub-vector pointer (utilities.cpp:31-58)");
  cudaStream_t s = (cudaStream_t)stream;
  EB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), s));
  eb::RhsParams P = make_params(c, w, wdot);
  if (slow_mode) {
    P.slow_mode = 1;
    P.inv_energy_units = 1.0 / energy_units;
    P.et_rw = const_cast<double*>(w[4]);
  }
  const long n[3] = {P.nx, P.ny, P.nz};
  // interior box and boundary shells (host_setup.h: overlap_boxes), cut along the tile grid of a launch over
  // the whole box
  eb::BoxList B;
  B.count = 0;
  bool interior = false;
  if (c->any_remote) {
    const long z0[3] = {0, 0, 0};
    const int vi0 = c->variant_part[0];
    const eb::LaunchGeom L0 = eb::launch_geom(z0, n, 5 + P.nchem, kVariants[vi0].threads, c->pair_sync, c->ctas_target,
                                              c->xc && kVariants[vi0].fnx[0][0] != nullptr);
    const long pitch[3] = {L0.xc ? L0.tx : L0.tx - 1, L0.ty - 1, L0.seg_len};
    bool rem[6];
    for (int f = 0; f < 6; f++) rem[f] = c->remote[f];
    interior = eb::overlap_boxes(n, rem, pitch, c->thick_shells != 0 && L0.tx == 32 && L0.ty == kVariants[vi0].threads / 32, &B);
  }
  int rc;
  if (c->any_remote && (!c->overlap || !interior)) {
    // Exchange first: pack, send / receive the halo slabs while the pre-pass runs, wait, then ONE launch
    // over the whole box (boundary tiles read the slabs).  The slabs of a 512^3 box are 94 MB per face
    // -- a few hundred microseconds over NVLink -- against a 60 ms kernel, so hiding them behind an
    // interior launch buys less than the six shell launches and the smaller interior grid cost
    // (profiles/README.md, 8 GPUs).  EULERB200_OVERLAP=1 restores the interior / shell schedule.
    // (slow mode: the pre-pass rebuilds the total energy the pack reads, so it goes first)
    c->prof_pack_first = !slow_mode;
    EB_PREC(c, 0, s);
    if (slow_mode && (rc = launch_aux(c, P, 0, P.nz, s))) return rc;
    if (slow_mode) EB_PREC(c, 1, s);
    if ((rc = exchange_start(c, w, s))) return rc;
    for (int f = 0; f < 6; f++) eb::ghost_face(c->cfg, f, c->recv_cur[f], &P.ghost[f]);   // slabs of this exchange
    if (!slow_mode) EB_PREC(c, 1, s);
    if (!slow_mode && (rc = launch_aux(c, P, 0, P.nz, s))) return rc;
    EB_PREC(c, 2, s);
    EB_PREC(c, 3, s);
    if ((rc = exchange_end(c, s))) return rc;
    EB_PREC(c, 4, s);
    const long z[3] = {0, 0, 0};
    if ((rc = launch_box(c, P, z, n, s))) return rc;
    EB_PREC(c, 5, s);
    return profile_collect(c);
  }
  c->prof_pack_first = false;
  EB_PREC(c, 0, s);
  // the exchange is started before the pre-pass is queued, so that its pack kernels (on the exchange stream)
  // do not wait for it (slow mode: the pre-pass rebuilds the total energy the pack reads, so it goes first)
  const bool early_start = c->any_remote && !slow_mode;
  if (early_start && (rc = exchange_start(c, w, s))) return rc;
  {
    int rc_ = launch_aux(c, P, 0, P.nz, s);
    if (rc_) return rc_;
  }
  EB_PREC(c, 1, s);
  if (!c->any_remote) {
    const long lo0[3] = {0, 0, 0};
    int rc_ = launch_box(c, P, lo0, n, s);
    if (rc_) return rc_;
    EB_PREC(c, 5, s);
    return profile_collect(c);
  }
  // Overlap (the structure of utilities.cpp:61 -> 76-116 -> 119 -> 123-195): start the
  // exchange, evaluate every cell whose stencils stay clear of the remote faces, wait for
  // the halos, then evaluate the remaining shell as non-overlapping slabs.
  if (!early_start && (rc = exchange_start(c, w, s))) return rc;
  for (int f = 0; f < 6; f++) eb::ghost_face(c->cfg, f, c->recv_cur[f], &P.ghost[f]);   // slabs of this exchange
  EB_PREC(c, 2, s);
  const long* lo = B.lo[0];
  const long* hi = B.hi[0];
  if (c->overlap == 2) {
    // Early shells: the slabs do not depend on the interior launch, only on the pre-pass (per-cell arrays)
    // and on the halo.  They go to three high-priority streams that wait for exactly those two, so their
    // CTAs slip into the SMs as interior CTAs retire -- a few milliseconds into the interior launch --
    // instead of running as a sparsely filled tail after it.
    EB_CUDA(c, cudaEventRecord(c->ev_fork, s));                      // pre-pass (and the NCCL-path pack) done
    rc = launch_box(c, P, lo, hi, s);                                // interior first: it starts at once
    if (rc) return rc;
    EB_PREC(c, 3, s);
    EB_PREC(c, 4, s);
    for (int h = 0; h < 3; h++) EB_CUDA(c, cudaStreamWaitEvent(c->slab_stream[h], c->ev_fork, 0));
    rc = exchange_end(c, c->slab_stream[0]);
    if (rc) return rc;
    EB_CUDA(c, cudaEventRecord(c->ev_halo, c->slab_stream[0]));
    for (int h = 1; h < 3; h++) EB_CUDA(c, cudaStreamWaitEvent(c->slab_stream[h], c->ev_halo, 0));
    for (int b = 1; b < B.count; b++)
      if ((rc = launch_box(c, P, B.lo[b], B.hi[b], c->slab_stream[(b - 1) % 3]))) return rc;
    for (int h = 0; h < 3; h++) {
      EB_CUDA(c, cudaEventRecord(c->ev_join[h], c->slab_stream[h]));
      EB_CUDA(c, cudaStreamWaitEvent(s, c->ev_join[h], 0));
    }
    EB_PREC(c, 5, s);
    return profile_collect(c);
  }
  rc = launch_box(c, P, lo, hi, s);
  if (rc) return rc;
  EB_PREC(c, 3, s);
  rc = exchange_end(c, s);
  if (rc) return rc;
  EB_PREC(c, 4, s);
  {
    // the slabs are independent and individually too small to fill the GPU: spread them over the
    // compute stream and two helper streams, then join
    cudaStream_t lane[3] = {s, c->slab_stream[0], c->slab_stream[1]};
    EB_CUDA(c, cudaEventRecord(c->ev_fork, s));
    for (int h = 0; h < 2; h++) EB_CUDA(c, cudaStreamWaitEvent(c->slab_stream[h], c->ev_fork, 0));
    for (int b = 1; b < B.count; b++)
      if ((rc = launch_box(c, P, B.lo[b], B.hi[b], lane[(b - 1) % 3]))) return rc;
    for (int h = 0; h < 2; h++) {
      EB_CUDA(c, cudaEventRecord(c->ev_join[h], c->slab_stream[h]));
      EB_CUDA(c, cudaStreamWaitEvent(s, c->ev_join[h], 0));
    }
  }
  EB_PREC(c, 5, s);
  return profile_collect(c);
}

int eulerb200_state_flag(eulerb200_ctx* c, void* stream, int32_t* bits)
{
  if (!c || !bits) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  EB_CUDA(c, cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  EB_CUDA(c, cudaStreamSynchronize(s));
  *bits = *c->h_flag;
  return 0;
}

int eulerb200_rhs(eulerb200_ctx* c, double t, const double* const* w, double* const* wdot, void* stream)
{
  int rc = eulerb200_rhs_async(c, t, w, wdot, stream);
  if (rc) return rc;
  int32_t bits = 0;
  rc = eulerb200_state_flag(c, stream, &bits);
  if (rc) return rc;
  return decode_flag(c, bits);
}

int eulerb200_rhs_host(eulerb200_ctx* c, double t, const double* const* wh, double* const* wdh)
{
  if (!c || !wh || !wdh) return -1;
  const eulerb200_config& g = c->cfg;
  const long plane = g.nxl * g.nyl, N = plane * g.nzl;
  const int nsub = 5 + (g.nchem > 0 ? 1 : 0);
  EB_CUDA(c, cudaSetDevice(c->device));
  { int rc_ = ensure_staging(c); if (rc_) return rc_; }
  if (c->any_remote) {
    // with remote neighbours the halo exchange needs the whole state: no slab pipeline
    for (int f = 0; f < nsub; f++) {
      EB_CUDA(c, cudaMemcpyAsync(c->stage_w[f], wh[f], sizeof(double) * N * (f < 5 ? 1 : g.nchem), cudaMemcpyHostToDevice, c->s_cmp));
      if (c->forcing_in_wdot)      // the hook's G travels with the state
        EB_CUDA(c, cudaMemcpyAsync(c->stage_wdot[f], wdh[f], sizeof(double) * N * (f < 5 ? 1 : g.nchem), cudaMemcpyHostToDevice, c->s_cmp));
    }
    int rc = eulerb200_rhs_async(c, t, c->stage_w, c->stage_wdot, c->s_cmp);
    if (rc) return rc;
    for (int f = 0; f < nsub; f++)
      EB_CUDA(c, cudaMemcpyAsync(wdh[f], c->stage_wdot[f], sizeof(double) * N * (f < 5 ? 1 : g.nchem), cudaMemcpyDeviceToHost, c->s_cmp));
  } else {
    // z-slab pipeline: upload slab s+1 while slab s is evaluated and slab s-1 is downloaded.
    // A slab can be evaluated once the slab above it is resident (3-plane stencil reach);
    // with a periodic wrap in z the first slab also needs the last one, so it goes last.
    int S = (int)std::min<long>(eulerb200_ctx::kMaxSlabs, std::max<long>(1, g.nzl / 8));
    if (const char* ev = getenv("EULERB200_HOST_SLABS")) S = std::max(1, std::min(S, atoi(ev)));
    long zb[eulerb200_ctx::kMaxSlabs + 1];
    for (int s = 0; s <= S; s++) zb[s] = g.nzl * s / S;
    EB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->s_cmp));
    for (int s = 0; s < S; s++) {
      for (int f = 0; f < nsub; f++) {
        const long m = (f < 5 ? 1 : g.nchem);
        EB_CUDA(c, cudaMemcpyAsync(c->stage_w[f] + zb[s] * plane * m, wh[f] + zb[s] * plane * m,
                                   sizeof(double) * (zb[s + 1] - zb[s]) * plane * m, cudaMemcpyHostToDevice, c->s_h2d));
        if (c->forcing_in_wdot)    // the hook's G travels with the state
          EB_CUDA(c, cudaMemcpyAsync(c->stage_wdot[f] + zb[s] * plane * m, wdh[f] + zb[s] * plane * m,
                                     sizeof(double) * (zb[s + 1] - zb[s]) * plane * m, cudaMemcpyHostToDevice, c->s_h2d));
      }
      EB_CUDA(c, cudaEventRecord(c->ev_up[s], c->s_h2d));
    }
    const bool wrap = (g.nbr[4] == g.rank);
    const eb::RhsParams P = make_params(c, c->stage_w, c->stage_wdot);
    int aux_done = 0;     // slabs whose per-cell derived values are computed
    for (int q = 0; q < S; q++) {
      const int s = wrap ? (q + 1) % S : q;
      const int need = (wrap && s == 0) ? S - 1 : std::min(s + 1, S - 1);
      EB_CUDA(c, cudaStreamWaitEvent(c->s_cmp, c->ev_up[need], 0));
      for (; aux_done <= need; aux_done++) {
        int rc_ = launch_aux(c, P, zb[aux_done], zb[aux_done + 1], c->s_cmp);
        if (rc_) return rc_;
      }
      const long lo[3] = {0, 0, zb[s]}, hi[3] = {g.nxl, g.nyl, zb[s + 1]};
      int rc = launch_box(c, P, lo, hi, c->s_cmp);
      if (rc) return rc;
      EB_CUDA(c, cudaEventRecord(c->ev_done[s], c->s_cmp));
      EB_CUDA(c, cudaStreamWaitEvent(c->s_d2h, c->ev_done[s], 0));
      for (int f = 0; f < nsub; f++) {
        const long m = (f < 5 ? 1 : g.nchem);
        EB_CUDA(c, cudaMemcpyAsync(wdh[f] + zb[s] * plane * m, c->stage_wdot[f] + zb[s] * plane * m,
                                   sizeof(double) * (zb[s + 1] - zb[s]) * plane * m, cudaMemcpyDeviceToHost, c->s_d2h));
      }
    }
    EB_CUDA(c, cudaStreamSynchronize(c->s_d2h));
  }
  int32_t bits = 0;
  int rc = eulerb200_state_flag(c, c->s_cmp, &bits);
  if (rc) return rc;
  return decode_flag(c, bits);
}

int eulerb200_rhs_any(eulerb200_ctx* c, double t, const double* const* w, double* const* wdot, void* stream)
{
  if (!c || !w || !wdot || !w[0]) return -1;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, w[0]);
  if (e != cudaSuccess) { cudaGetLastError(); return eulerb200_rhs_host(c, t, w, wdot); }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return eulerb200_rhs(c, t, w, wdot, stream);
  return eulerb200_rhs_host(c, t, w, wdot);
}

int eulerb200_stability(eulerb200_ctx* c, const double* const* w, double cfl, double* dt_stab, void* stream)
{
  if (!c || !w || !dt_stab) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const long N = c->cfg.nxl * c->cfg.nyl * c->cfg.nzl;
  EB_CUDA(c, cudaMemsetAsync(c->d_alpha, 0, sizeof(unsigned long long), s));
  const unsigned blocks = (unsigned)std::min<long>((N + 255) / 256, 148L * 8);
  eb::wavespeed_kernel<<<blocks, 256, 0, s>>>(w[0], w[1], w[2], w[3], w[4], N, c->cfg.gamma, c->d_alpha);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  if (c->cfg.nranks > 1 && c->comm)   // utilities.cpp:516
    EB_NCCL(c, nccl().AllReduce(c->d_alpha, c->d_alpha, 1, ncclDouble, ncclMax, c->comm, s));
  EB_CUDA(c, cudaMemcpyAsync(c->h_alpha, c->d_alpha, sizeof(double), cudaMemcpyDeviceToHost, s));
  EB_CUDA(c, cudaStreamSynchronize(s));
  const double h = std::min(std::min(c->cfg.dx, c->cfg.dy), c->cfg.dz);
  *dt_stab = cfl * h / *c->h_alpha;   // utilities.cpp:520
  return 0;
}

int eulerb200_stability_any(eulerb200_ctx* c, const double* const* w, double cfl, double* dt_stab, void* stream)
{
  if (!c || !w || !w[0]) return -1;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, w[0]);
  if (e != cudaSuccess) cudaGetLastError();
  if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged))
    return eulerb200_stability(c, w, cfl, dt_stab, stream);
  int rc = ensure_staging(c);
  if (rc) return rc;
  const long N = c->cfg.nxl * c->cfg.nyl * c->cfg.nzl;
  for (int f = 0; f < 5; f++)
    EB_CUDA(c, cudaMemcpyAsync(c->stage_w[f], w[f], sizeof(double) * N, cudaMemcpyHostToDevice, c->s_cmp));
  return eulerb200_stability(c, c->stage_w, cfl, dt_stab, c->s_cmp);
}

int eulerb200_vec_lincomb(eulerb200_ctx* c, int32_t nterms, const double* coef, const double* const* x,
                          double* out, int64_t n, void* stream)
{
  if (!c || !coef || !x || !out || nterms < 1 || nterms > 16) return -1;
  eb::LinCombArgs a;
  a.nterms = nterms;
  for (int t = 0; t < nterms; t++) { a.c[t] = coef[t]; a.x[t] = x[t]; }
  const unsigned blocks = (unsigned)std::min<long>((n + 255) / 256, 148L * 16);
  eb::lincomb_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, out, n);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  return 0;
}

int eulerb200_vec_wrms_accum(eulerb200_ctx* c, const double* x, const double* y, double rtol, double atol,
                             int64_t n, double* acc, void* stream)
{
  if (!c || !x || !y || !acc) return -1;
  const unsigned blocks = (unsigned)std::min<long>((n + 255) / 256, 148L * 8);
  eb::wrms_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, rtol, atol, n, acc);
  c->launches++;
  EB_CUDA(c, cudaGetLastError());
  return 0;
}

// ||x||_WRMS over the whole ManyVector and all ranks: sqrt( sum (x_i/(rtol|y_i|+atol))^2 / nglobal )
int eulerb200_vec_wrms(eulerb200_ctx* c, const double* const* x, const double* const* y, double rtol, double atol,
                       int64_t nglobal, double* result, void* stream)
{
  if (!c || !x || !y || !result || nglobal <= 0) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = reinterpret_cast<double*>(c->d_alpha);      // 8-byte device scratch shared with stability()
  EB_CUDA(c, cudaMemsetAsync(acc, 0, sizeof(double), s));
  const long N = c->cfg.nxl * c->cfg.nyl * c->cfg.nzl;
  for (int f = 0; f < 5 + (c->cfg.nchem > 0 ? 1 : 0); f++) {
    int rc = eulerb200_vec_wrms_accum(c, x[f], y[f], rtol, atol, f < 5 ? N : N * c->cfg.nchem, acc, stream);
    if (rc) return rc;
  }
  if (c->cfg.nranks > 1 && c->comm)
    EB_NCCL(c, nccl().AllReduce(acc, acc, 1, ncclDouble, ncclSum, c->comm, s));
  EB_CUDA(c, cudaMemcpyAsync(c->h_alpha, acc, sizeof(double), cudaMemcpyDeviceToHost, s));
  EB_CUDA(c, cudaStreamSynchronize(s));
  *result = sqrt(*c->h_alpha / (double)nglobal);
  return 0;
}

// Plain device-memory helpers so that a C/C++ host driver needs no CUDA headers.
void* eulerb200_device_alloc(int64_t bytes)
{
  void* p = nullptr;
  if (bytes <= 0 || cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void eulerb200_device_free(void* p) { if (p) cudaFree(p); }
void* eulerb200_managed_alloc(int64_t bytes)
{
  void* p = nullptr;
  if (bytes <= 0 || cudaMallocManaged(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
int eulerb200_synchronize(eulerb200_ctx* c)
{
  if (!c) return -1;
  EB_CUDA(c, cudaSetDevice(c->device));
  EB_CUDA(c, cudaDeviceSynchronize());
  return 0;
}
int eulerb200_copy_to_device(void* dst, const void* src, int64_t bytes)
{
  return cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : fail(nullptr, -2, "cudaMemcpy H2D failed");
}
int eulerb200_copy_to_host(void* dst, const void* src, int64_t bytes)
{
  return cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : fail(nullptr, -2, "cudaMemcpy D2H failed");
}

int64_t eulerb200_launch_count(const eulerb200_ctx* c) { return c ? c->launches : -1; }

int eulerb200_profile(eulerb200_ctx* c, int32_t on, int32_t reset, double* out)
{
  if (!c) return -1;
  EB_CUDA(c, cudaSetDevice(c->device));
  if (out) {
    const double nn = c->pacc_n > 0 ? c->pacc_n : 1.0;
    for (int q = 0; q < 7; q++) out[q] = c->pacc[q] / nn;
    out[7] = c->pacc_n;
  }
  if (reset) { for (int q = 0; q < 7; q++) c->pacc[q] = 0; c->pacc_n = 0; }
  if (on && !c->pev[0])
    for (int q = 0; q < 8; q++) EB_CUDA(c, cudaEventCreate(&c->pev[q]));
  c->profile_on = on != 0;
  return 0;
}

int eulerb200_set_forcing_in_wdot(eulerb200_ctx* c, int32_t on)
{
  if (!c) return -1;
  c->forcing_in_wdot = (on != 0);
  return 0;
}

int eulerb200_fp64_peak(double* tflops)
{
  if (!tflops) return -1;
  double* d = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) return fail(nullptr, -2, "cudaMalloc failed");
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 20000;
  dfma_peak_kernel<<<blocks, threads>>>(d, 2000, 0.999999, 1e-9);   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return fail(nullptr, -2, "dfma kernel failed"); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return 0;
}

}  // extern "C"
