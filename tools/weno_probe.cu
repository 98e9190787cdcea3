// Micro-probe (tuning aid, not product): how busy can the FP64 pipe get on the WENO arithmetic
// itself (registers only, no memory) as a function of threads per SM and reconstructions in
// flight per thread?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sundials-manyvector-demo_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "euler_math.cuh"

template <int NPAR>
__global__ void weno_loop(double* out, int iters, double seed)
{
  double v[NPAR][6];
#pragma unroll
  for (int p = 0; p < NPAR; p++)
#pragma unroll
    for (int l = 0; l < 6; l++) v[p][l] = seed + 0.01 * (threadIdx.x % 7) + 0.1 * l + p;
  double up[6], um[6];
#pragma unroll
  for (int l = 0; l < 6; l++) { up[l] = 1.5 + 0.01 * l; um[l] = -0.5 + 0.01 * l; }
  double acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int p = 0; p < NPAR; p++) {
      const double f = eb::tracer_face(v[p], up, um);   // 2 reconstructions
#pragma unroll
      for (int l = 0; l < 6; l++) v[p][l] = fma(v[p][l], 0.9999 + 1e-5 * l, 1e-9 * f);   // all inputs change
      acc += f;
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

template <int NPAR>
double run(int threads, int sms, double* d)
{
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  weno_loop<NPAR><<<sms, threads>>>(d, 100, 1.0);
  cudaEventRecord(e0);
  weno_loop<NPAR><<<sms, threads>>>(d, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // FP64-pipe instructions per tracer_face: 2 x 43 (weno5) + 10 (products) + 2
  const double lane_instr = (98.0 + 7.0) * NPAR * iters * (double)sms * threads;
  return lane_instr / (ms * 1e-3) / 16.85e12;     // fraction of the measured DFMA issue rate (33.7 TF / 2)
}

int main()
{
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d;
  cudaMalloc(&d, 8);
  printf("FP64 pipe utilisation on tracer_face (no memory), one CTA per SM\nthreads  1 in flight  2 in flight  4 in flight\n");
  const int ts[] = {128, 256, 384, 512};
  for (int t : ts) printf("%7d  %10.2f  %11.2f  %11.2f\n", t, run<1>(t, sms, d), run<2>(t, sms, d), run<4>(t, sms, d));
  return 0;
}
