// Tuning aid (not product): time the RHS kernels of rhs_kernel.cuh on a synthetic 512^3-like state
// with measurement-only ablations compiled in (-DEB_ABLATE=<mask>, see rhs_kernel.cuh), to attribute
// kernel time to barriers / loads / stores.  Results of ablated runs are wrong by construction.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DEB_ABLATE=<m> -I sundials-manyvector-demo_b200/csrc -o ablate_<m> tools/ablate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define EB_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#include "host_setup.h"

__global__ void fill_state(double* rho, double* mx, double* my, double* mz, double* et, double* chem, long N, int nchem)
{
  for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (long)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)c * 6364136223846793005ull + 1442695040888963407ull;
    auto U = [&]() { h ^= h >> 29; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 32; return (double)(h >> 11) * (1.0 / 9007199254740992.0); };
    const double r = 1.0 + 0.5 * U(), vx = 0.3 * (U() - 0.5), vy = 0.3 * (U() - 0.5), vz = 0.3 * (U() - 0.5), p = 1.0 + 0.5 * U();
    rho[c] = r; mx[c] = r * vx; my[c] = r * vy; mz[c] = r * vz;
    et[c] = p / (5.0 / 3.0 - 1.0) + 0.5 * r * (vx * vx + vy * vy + vz * vz);
    for (int v = 0; v < nchem; v++) chem[c * nchem + v] = U();
  }
}

template <int T, int PART>
float run(eb::RhsParams P, int pair, int reps, const char* tag)
{
  const int nf = PART == eb::PART_ALL ? 5 + P.nchem : (PART == eb::PART_FLUID ? 5 : P.nchem);
  const eb::LaunchGeom L = eb::launch_geom(P.lo, P.hi, nf, T, pair, 5920);
  P.seg_len = L.seg_len; P.pair_sync = L.pair;
  auto fn = eb::rhs_fused_kernel<T, 1, false, false, PART, T / 32>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
  const int pct = (int)std::min<size_t>(100, (100 * (L.smem + 1024) + 228 * 1024 - 1) / (228 * 1024));
  cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fn<<<dim3(L.gx, L.gy, L.gz), dim3(L.tx, L.ty, 1), L.smem>>>(P);
  cudaEventRecord(e0);
  for (int r = 0; r < reps; r++) fn<<<dim3(L.gx, L.gy, L.gz), dim3(L.tx, L.ty, 1), L.smem>>>(P);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("ablate=%2d %-22s threads=%d pair=%d smem=%zu  %8.3f ms  %s\n", EB_ABLATE, tag, T, L.pair, L.smem, ms / reps,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
  return ms / reps;
}

int main(int argc, char** argv)
{
  const long n = argc > 1 ? atol(argv[1]) : 512;
  const int nchem = argc > 2 ? atoi(argv[2]) : 10;
  const long N = n * n * n;
  double *w[6], *wd[6], *aux[4], *chemT;
  for (int f = 0; f < 5; f++) { cudaMalloc(&w[f], 8 * N); cudaMalloc(&wd[f], 8 * N); }
  cudaMalloc(&w[5], 8 * N * nchem); cudaMalloc(&wd[5], 8 * N * nchem);
  for (int q = 0; q < 4; q++) cudaMalloc(&aux[q], 8 * N);
  cudaMalloc(&chemT, 8 * N * 2 * ((nchem + 1) / 2));
  int* flag; cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
  fill_state<<<148 * 8, 256>>>(w[0], w[1], w[2], w[3], w[4], w[5], N, nchem);
  eulerb200_config cfg = {};
  cfg.nxl = cfg.nyl = cfg.nzl = n; cfg.nchem = nchem; cfg.dx = cfg.dy = cfg.dz = 1.0 / n; cfg.gamma = 5.0 / 3.0;
  for (int f = 0; f < 6; f++) { cfg.bc[f] = EULERB200_BC_REFLECTING; cfg.nbr[f] = EULERB200_NO_NEIGHBOR; }
  eb::RhsParams P;
  P.nx = P.ny = P.nz = n; P.nchem = nchem; P.gamma = cfg.gamma; P.rdx = P.rdy = P.rdz = EB_RD_SCALE * (double)n; P.dx = P.dy = P.dz = 1.0 / n;
  for (int f = 0; f < 5; f++) P.forcing[f] = 0.0;
  for (int f = 0; f < 6; f++) { P.w[f] = w[f]; P.wdot[f] = wd[f]; eb::ghost_face(cfg, f, nullptr, &P.ghost[f]); }
  for (int q = 0; q < 4; q++) P.aux[q] = aux[q];
  P.chemT = (argc > 3 && atoi(argv[3]) == 0) ? nullptr : chemT;
  P.state_flag = flag; P.pair_sync = 0; P.vec_store = 1; P.slow_mode = 0; P.inv_energy_units = 1.0; P.et_rw = nullptr;
  for (int d = 0; d < 3; d++) { P.lo[d] = 0; P.hi[d] = n; }
  P.seg_len = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  eb::aux_kernel<<<148 * 16, 256>>>(P, aux[0], aux[1], aux[2], aux[3], chemT, 0, N);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("aux_kernel (with species transposition) %.3f ms\n", ms);
  const int reps = 3;
#ifndef ABL_QUICK
  run<384, eb::PART_ALL>(P, 2, reps, "fused");
  run<384, eb::PART_FLUID>(P, 2, reps, "fluid");
  run<384, eb::PART_TRACERS>(P, 2, reps, "species");
  run<640, eb::PART_TRACERS>(P, 0, reps, "species");
#endif
#ifdef ABL_T
  run<ABL_T, eb::ABL_PART>(P, 2, reps, "quick");
#elif !defined(ABL_FUSED_ONLY)
  run<512, eb::PART_TRACERS>(P, 2, reps, "species");
#else
  run<384, eb::PART_ALL>(P, 2, reps, "fused");
#endif
  return 0;
}
