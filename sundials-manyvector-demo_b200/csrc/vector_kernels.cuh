// ---------------------------------------------------------------------------
// vector_kernels.cuh -- the one-pass, HBM-bound kernels around the RHS (sm_100a):
//   wavespeed_kernel   local part of `stability` (utilities.cpp:505-513)
//   lincomb_kernel     stage combinations of the explicit driver loop (N_VLinearCombination)
//   wrms_kernel        weighted RMS norm of the error test (N_VWrmsNorm)
// Warp-shuffle reductions, one atomic per CTA.  Like the other kernel headers the file also
// compiles under g++ with tests/emu/cuda_emu.h (shuffles and atomics emulated) for the CPU tier.
// ---------------------------------------------------------------------------
#pragma once
#include "euler_math.cuh"

namespace eb {

#if defined(__CUDACC__) || defined(EB_CUDA_EMU)

// utilities.cpp:505-513: alpha = max | |mx/rho| + sqrt(gamma p / rho) |  (my, mz only via p).
// Warp-shuffle then one atomic per CTA; non-negative doubles order like their bit patterns.
__global__ void wavespeed_kernel(const double* __restrict__ rho, const double* __restrict__ mx,
                                 const double* __restrict__ my, const double* __restrict__ mz,
                                 const double* __restrict__ et, long N, double gamma,
                                 unsigned long long* __restrict__ out)
{
  double alpha = 0.0;
  for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < N; c += (long)gridDim.x * blockDim.x) {
    const double r = rho[c], a = mx[c], b = my[c], d = mz[c];
    const double u = fabs(a / r);
    const double p = (gamma - 1.0) * (et[c] - (a * a + b * b + d * d) * 0.5 / r);
    const double x = fabs(u + sun_sqrt(gamma * p / r));
    alpha = (alpha < x) ? x : alpha;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double y = __shfl_xor_sync(0xffffffffu, alpha, o);
    alpha = (alpha < y) ? y : alpha;
  }
  __shared__ double part[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = alpha;
  __syncthreads();
  if (wid == 0) {
    alpha = (lane < (blockDim.x >> 5)) ? part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double y = __shfl_xor_sync(0xffffffffu, alpha, o);
      alpha = (alpha < y) ? y : alpha;
    }
    if (lane == 0) atomicMax(out, (unsigned long long)__double_as_longlong(alpha));
  }
}

// ---- vector operations of the explicit driver loop: the stage
// combinations and the weighted RMS norm ARKODE evaluates through N_VLinearCombination /
// N_VWrmsNorm on the MPIManyVector.  One pass each, HBM bound, grid-stride over a grid that
// is a multiple of the SM count.
struct LinCombArgs {
  int nterms;
  double c[16];
  const double* x[16];
};
__global__ void lincomb_kernel(const LinCombArgs a, double* __restrict__ out, long n)
{
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double s = a.c[0] * a.x[0][i];
#pragma unroll 1
    for (int t = 1; t < a.nterms; t++) s = fma(a.c[t], a.x[t][i], s);
    out[i] = s;
  }
}
// sum_i (x_i / (rtol*|y_i| + atol))^2  accumulated into *acc (one atomicAdd per CTA)
__global__ void wrms_kernel(const double* __restrict__ x, const double* __restrict__ y, double rtol, double atol,
                            long n, double* __restrict__ acc)
{
  double s = 0.0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double q = x[i] / fma(rtol, fabs(y[i]), atol);
    s = fma(q, q, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = s;
  __syncthreads();
  if (wid == 0) {
    s = (lane < (blockDim.x >> 5)) ? part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(acc, s);
  }
}

#endif  // __CUDACC__ || EB_CUDA_EMU

}  // namespace eb
