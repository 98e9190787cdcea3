// ---------------------------------------------------------------------------
// cuda_emu.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE, never linked into the product.
//
// Just enough of the CUDA execution model to run the kernel SOURCE of
// sundials-manyvector-demo_b200/csrc/rhs_kernel.cuh on the CPU in the "not gpu" test
// tier, so that index arithmetic, ghost maps, shared-memory exchange and barrier
// placement can be checked against the oracle in a container without a GPU.
// It proves nothing about performance and is not a fallback: the product library
// (libeulerb200.so) contains no host path and fails loudly without a device.
//
// Model: one CTA at a time; each CUDA thread is a ucontext fiber.  Barriers (__syncthreads,
// the named "bar.sync id, count" the kernel uses between neighbouring warp rows, __syncwarp)
// count arrivals; a fiber that is not the last to arrive yields to a round-robin scheduler
// until the barrier's generation changes.  threadIdx/blockIdx/blockDim/gridDim are plain
// globals that the scheduler rewrites on every switch.
// ---------------------------------------------------------------------------
#pragma once
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define EB_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static            /* one CTA runs at a time */

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_emu { unsigned x, y, z; };
struct double2 { double x, y; };

namespace cuda_emu {
inline uint3_emu& tidx() { static uint3_emu v; return v; }
inline uint3_emu& bidx() { static uint3_emu v; return v; }
inline dim3& bdim() { static dim3 v; return v; }
inline dim3& gdim() { static dim3 v; return v; }
inline void*& shared_base() { static void* p = nullptr; return p; }

struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done; uint3_emu tid; };
inline std::vector<Fiber>& fibers() { static std::vector<Fiber> f; return f; }
inline int& current() { static int c = -1; return c; }
inline ucontext_t& sched_ctx() { static ucontext_t c; return c; }

inline void yield_to_scheduler() { swapcontext(&fibers()[current()].ctx, &sched_ctx()); }

// slots 0..15: the hardware's named barriers (0 = __syncthreads); 16 + w: __syncwarp of warp w
struct Barrier { int arrived; unsigned gen; };
inline std::vector<Barrier>& barriers() { static std::vector<Barrier> b; return b; }
inline void barrier_wait(int slot, int count)
{
  Barrier& b = barriers()[slot];
  const unsigned g = b.gen;
  if (++b.arrived >= count) { b.arrived = 0; b.gen++; return; }
  long spins = 0;
  while (barriers()[slot].gen == g) {
    if (++spins > 100000000L) { fprintf(stderr, "cuda_emu: barrier %d never completes (deadlock)\n", slot); abort(); }
    yield_to_scheduler();
  }
}

template <class Kernel, class Params> struct Launch {
  static Kernel kernel;
  static const Params* params;
  static void entry() { kernel(*params); fibers()[current()].done = true; swapcontext(&fibers()[current()].ctx, &sched_ctx()); }
};
template <class K, class P> K Launch<K, P>::kernel;
template <class K, class P> const P* Launch<K, P>::params;

// Run kernel(params) over grid x block with `shmem` bytes of dynamic shared memory.
template <class Params>
void launch(void (*kernel)(const Params), dim3 grid, dim3 block, size_t shmem, const Params& params)
{
  typedef Launch<void (*)(const Params), Params> L;
  L::kernel = kernel;
  L::params = &params;
  gdim() = grid; bdim() = block;
  const int T = block.x * block.y * block.z;
  std::vector<char> smem(shmem + 64);
  static std::vector<std::vector<char> > stacks;                             // reused by every CTA and launch
  while ((int)stacks.size() < T) stacks.push_back(std::vector<char>(96 * 1024));
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        bidx().x = bx; bidx().y = by; bidx().z = bz;
        memset(smem.data(), 0xFF, smem.size());       // poison: NaNs if read before written
        shared_base() = smem.data();
        fibers().assign(T, Fiber());
        barriers().assign(16 + (T + 31) / 32, Barrier{0, 0u});
        for (int t = 0; t < T; t++) {
          Fiber& f = fibers()[t];
          f.done = false;
          f.tid.x = t % block.x; f.tid.y = (t / block.x) % block.y; f.tid.z = t / (block.x * block.y);
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = stacks[t].data();
          f.ctx.uc_stack.ss_size = stacks[t].size();
          f.ctx.uc_link = &sched_ctx();
          makecontext(&f.ctx, (void (*)())L::entry, 0);
        }
        // each pass resumes every live fiber once, until it blocks in a barrier again or exits
        // (alternating the order between passes makes a missing barrier show up as wrong data)
        bool any = true;
        int pass = 0;
        while (any) {
          any = false;
          pass++;
          for (int q = 0; q < T; q++) {
            const int t = (pass & 1) ? q : T - 1 - q;
            if (fibers()[t].done) continue;
            any = true;
            current() = t;
            tidx() = fibers()[t].tid;
            swapcontext(&sched_ctx(), &fibers()[t].ctx);
          }
        }
      }
}
// Kernels without barriers or shared memory (grid-stride loops): every thread runs to completion in turn.
template <class K, class... A>
void launch_plain(K kernel, dim3 grid, dim3 block, A... args)
{
  gdim() = grid; bdim() = block;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        bidx().x = bx; bidx().y = by; bidx().z = bz;
        for (unsigned tz = 0; tz < block.z; tz++)
          for (unsigned ty = 0; ty < block.y; ty++)
            for (unsigned tx = 0; tx < block.x; tx++) {
              tidx().x = tx; tidx().y = ty; tidx().z = tz;
              kernel(args...);
            }
      }
}
template <class K, class A, class B, class C> void launch2(K k, dim3 g, dim3 b, A a0, B a1, C a2) { launch_plain(k, g, b, a0, a1, a2); }
template <class K, class A, class B, class C, class D> void launch3(K k, dim3 g, dim3 b, A a0, B a1, C a2, D a3) { launch_plain(k, g, b, a0, a1, a2, a3); }
// grid-stride "for each cell" helper for trivially parallel pre-pass kernels
template <class F> void launch_aux_like(F f, long n) { for (long c = 0; c < n; c++) f(c); }
}  // namespace cuda_emu

#define threadIdx (cuda_emu::tidx())
#define blockIdx (cuda_emu::bidx())
#define blockDim (cuda_emu::bdim())
#define gridDim (cuda_emu::gdim())
#define EB_DYN_SMEM(type, name) type* name = (type*)cuda_emu::shared_base()

inline int emu_block_threads() { return (int)(cuda_emu::bdim().x * cuda_emu::bdim().y * cuda_emu::bdim().z); }
inline void __syncthreads() { cuda_emu::barrier_wait(0, emu_block_threads()); }
inline void eb_bar_sync(int id, int count) { cuda_emu::barrier_wait(id, count); }
inline void __syncwarp()
{
  const int T = emu_block_threads(), w = cuda_emu::current() / 32;
  cuda_emu::barrier_wait(16 + w, (T - 32 * w < 32) ? T - 32 * w : 32);
}
// cp.async.bulk + mbarrier (rhs_fused_kernel<..., STAGE>): the copy happens at issue; the barrier's
// phase flips when the expected bytes have arrived; a waiting fibre yields until then.
// (arrive-count barriers of rhs_fused_kernel<..., XC>: the phase flips when `count` arrivals are in and no
// bytes are outstanding.)  Eight bytes, like the hardware's.
struct EmuMbar { unsigned char phase, count, arrived, pad; int pending; };
static_assert(sizeof(EmuMbar) == 8, "an mbarrier is one 64-bit word");
inline EmuMbar* emu_mbar(unsigned long long* bar) { return reinterpret_cast<EmuMbar*>(bar); }
inline void emu_mbar_check(EmuMbar* m) { if (m->arrived >= m->count && m->pending == 0) { m->phase ^= 1u; m->arrived = 0; } }
inline void eb_mbar_init(unsigned long long* bar, int count)
{
  EmuMbar* m = emu_mbar(bar);
  m->phase = 0; m->count = (unsigned char)count; m->arrived = 0; m->pad = 0; m->pending = 0;
}
inline void eb_mbar_expect_tx(unsigned long long* bar, unsigned bytes)      // (arrive.expect_tx)
{
  EmuMbar* m = emu_mbar(bar);
  m->pending += (int)bytes; m->arrived++;
  emu_mbar_check(m);
}
inline void eb_mbar_arrive(unsigned long long* bar) { EmuMbar* m = emu_mbar(bar); m->arrived++; emu_mbar_check(m); }
inline void eb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  memcpy(dst, src, bytes);
  EmuMbar* m = emu_mbar(bar);
  m->pending -= (int)bytes;
  emu_mbar_check(m);
}
inline void eb_mbar_wait(unsigned long long* bar, unsigned parity)
{
  long spins = 0;
  while (emu_mbar(bar)->phase == parity) {
    if (++spins > 100000000L) { fprintf(stderr, "cuda_emu: mbarrier never completes\n"); abort(); }
    cuda_emu::yield_to_scheduler();
  }
}
inline int atomicOr(int* addr, int v) { int old = *addr; *addr = old | v; return old; }
inline double atomicAdd(double* addr, double v) { const double old = *addr; *addr = old + v; return old; }
inline unsigned long long atomicMax(unsigned long long* addr, unsigned long long v)
{
  const unsigned long long old = *addr;
  if (v > old) *addr = v;
  return old;
}
inline long long __double_as_longlong(double x) { long long y; memcpy(&y, &x, sizeof y); return y; }
// Full-warp butterfly exchange: every lane of the (complete) warp deposits its value, meets the others
// at the warp barrier, reads its partner's, and meets them again before the slot is reused.
inline double __shfl_xor_sync(unsigned, double v, int lane_mask)
{
  static std::vector<double> slot;
  const int t = cuda_emu::current(), T = emu_block_threads();
  if ((int)slot.size() < T) slot.resize(T);
  slot[t] = v;
  __syncwarp();
  const double r = slot[(t & ~31) | ((t & 31) ^ lane_mask)];
  __syncwarp();
  return r;
}
