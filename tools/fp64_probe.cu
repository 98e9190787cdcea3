// Micro-probe (tuning aid, not product): DFMA throughput on sm_100a as a function of resident
// warps per SM and independent chains per thread -> how much parallelism the FP64 pipe needs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma(double* out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r += x[i];
  if (r == 12345.678) out[0] = r;
}

template <int ILP>
double run(int warps_per_sm, int sms, double* d)
{
  const int threads = 32 * warps_per_sm, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma<ILP><<<sms, threads>>>(d, 1000, 0.999999, 1e-9);
  cudaEventRecord(e0);
  dfma<ILP><<<sms, threads>>>(d, iters, 0.999999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return 2.0 * ILP * iters * (double)sms * threads / (ms * 1e-3) / 1e12;
}

int main()
{
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d;
  cudaMalloc(&d, 8);
  printf("DFMA TFLOP/s on %d SMs (one CTA per SM)\nwarps/SM  ILP1    ILP2    ILP4    ILP8\n", sms);
  const int ws[] = {4, 8, 12, 16, 24, 32};
  for (int w : ws)
    printf("%7d  %6.2f  %6.2f  %6.2f  %6.2f\n", w, run<1>(w, sms, d), run<2>(w, sms, d), run<4>(w, sms, d), run<8>(w, sms, d));
  return 0;
}
