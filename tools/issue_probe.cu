// Tuning aid (not product): does a non-FP64 instruction issue "for free" next to a DFMA stream on B200?
// 8 independent DFMA chains per thread plus K independent integer (IMAD / LOP3) or FP32 (FFMA) instructions per
// iteration; time per iteration vs K tells whether the cost model is max(2 N_fp64, N_total) or 2 N_fp64 + N_other.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void probe(double* out, int iters, double a, double b, int ia, float fa)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  int n[16];
  float f[16];
#pragma unroll
  for (int q = 0; q < 16; q++) { n[q] = threadIdx.x + q; f[q] = threadIdx.x + q; }
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
#pragma unroll
    for (int q = 0; q < K; q++) {
      if (KIND == 0) n[q % 16] = n[q % 16] * ia + i;          // IMAD
      else if (KIND == 1) n[q % 16] = (n[q % 16] ^ ia) + 3;   // LOP3 / IADD3 (alu pipe)
      else f[q % 16] = fmaf(f[q % 16], fa, 1.0f);             // FFMA
    }
  }
  double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  int s = 0; float t = 0;
#pragma unroll
  for (int q = 0; q < 16; q++) { s += n[q]; t += f[q]; }
  if (r == 12345.678 || s == 123456789 || t == 1.2345f) out[0] = r + s + t;
}

template <int K, int KIND>
void run(double* d, int sms, int threads, int ctas)
{
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<K, KIND><<<sms * ctas, threads>>>(d, 200, 0.999999, 1e-9, 3, 0.999f);
  cudaEventRecord(e0);
  probe<K, KIND><<<sms * ctas, threads>>>(d, iters, 0.999999, 1e-9, 3, 0.999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // cycles per iteration per SMSP: warps per SMSP = threads*ctas/128
  const double wps = threads * ctas / 128.0;
  const double cyc = ms * 1e-3 * 1.965e9 / iters / wps;
  printf("kind=%d K=%2d threads/SM=%4d  %.3f ms  %.2f issue cycles per warp-iteration (8 DFMA + %d other; 16 = DFMA bound)\n",
         KIND, K, threads * ctas, ms, cyc, K);
}

int main()
{
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d; cudaMalloc(&d, 8);
  for (int cfg = 0; cfg < 2; cfg++) {
    const int threads = cfg == 0 ? 384 : 256, ctas = cfg == 0 ? 1 : 4;
    run<0, 0>(d, sms, threads, ctas); run<2, 0>(d, sms, threads, ctas); run<4, 0>(d, sms, threads, ctas); run<8, 0>(d, sms, threads, ctas); run<16, 0>(d, sms, threads, ctas);
    run<4, 1>(d, sms, threads, ctas); run<8, 1>(d, sms, threads, ctas); run<16, 1>(d, sms, threads, ctas);
    run<4, 2>(d, sms, threads, ctas); run<8, 2>(d, sms, threads, ctas); run<16, 2>(d, sms, threads, ctas);
  }
  return 0;
}
