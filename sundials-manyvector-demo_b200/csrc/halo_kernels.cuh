// ---------------------------------------------------------------------------
// halo_kernels.cuh -- the two face kernels of the halo exchange (sm_100a; O(surface), HBM bound):
//   pack_face_kernel / pack_face_warp_kernel
//                      ExchangeStart's send-buffer packing (euler3D.hpp:644-786); the destination
//                      is the local send slab (NCCL transport) or the neighbour's ghost slab
//                      (peer-store transport)
//   ghost_face_kernel  a face's ghost layers in the reference's receive-buffer layout, whatever
//                      their source (halo slab, periodic wrap, boundary-condition fill
//                      euler3D.hpp:797-1166): tests and the drop-in only
// Like rhs_kernel.cuh the file also compiles under g++ with tests/emu/cuda_emu.h, so that the wire
// layout is checked against the oracle in the CPU test tier.
// ---------------------------------------------------------------------------
#pragma once
#include "rhs_kernel.cuh"

namespace eb {

struct FaceGeom {
  long nx, ny, nz;
  int nchem, f;
  const double* w[6];
};

EB_HD void face_decode(const FaceGeom& g, long e, int& d, long& a, long& b, long& na)
{
  const int dir = g.f / 2;
  na = (dir == 0) ? g.ny : g.nx;
  d = (int)(e % 3);
  const long r = e / 3;
  a = r % na;
  b = r / na;
}
EB_HD long face_cell(const FaceGeom& g, long src, long a, long b)
{
  const int dir = g.f / 2;
  const long i = (dir == 0) ? src : a;
  const long j = (dir == 0) ? a : (dir == 1 ? src : b);
  const long k = (dir == 2) ? src : b;
  return i + g.nx * (j + g.ny * k);
}

// What this rank sends through face f: its three layers nearest that face in increasing
// index order, all NVAR values of a cell contiguous (euler3D.hpp:644-786).
#if defined(__CUDACC__) || defined(EB_CUDA_EMU)
// One thread per entry (cell of a layer): the reads of neighbouring threads are neighbouring cells.  For a
// local destination (the send slab of the NCCL transport): 0.11 ms per 512^2 face.
__global__ void pack_face_kernel(const FaceGeom g, double* __restrict__ buf, long nent)
{
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < nent; e += (long)gridDim.x * blockDim.x) {
  int d; long a, b, na;
  face_decode(g, e, d, a, b, na);
  const long n = (g.f / 2 == 0) ? g.nx : (g.f / 2 == 1 ? g.ny : g.nz);
  const long src = (g.f % 2 == 0) ? d : n - 3 + d;
  const long cell = face_cell(g, src, a, b);
  const int nv = 5 + g.nchem;
  double* o = buf + (long)nv * e;
#pragma unroll
  for (int v = 0; v < 5; v++) o[v] = g.w[v][cell];
  for (int v = 0; v < g.nchem; v++) o[5 + v] = g.w[5][cell * g.nchem + v];
  }
}

// The same with one WARP per entry (blocks of 32 x 8 threads; threadIdx.y picks the entry, the lanes its NVAR
// values), so that every store instruction writes one contiguous run of the destination -- for a destination
// across NVLink (the neighbour's ghost slab, peer-store transport), where the thread-per-entry kernel's
// NVAR scattered 8-byte remote stores per thread reach 38 GB/s (2.5 ms per face; this one 0.6 ms).
__global__ void pack_face_warp_kernel(const FaceGeom g, double* __restrict__ buf, long nent)
{
  const int nv = 5 + g.nchem;
  const long n = (g.f / 2 == 0) ? g.nx : (g.f / 2 == 1 ? g.ny : g.nz);
  for (long e = blockIdx.x * (long)blockDim.y + threadIdx.y; e < nent; e += (long)gridDim.x * blockDim.y) {
    int d; long a, b, na;
    face_decode(g, e, d, a, b, na);
    const long src = (g.f % 2 == 0) ? d : n - 3 + d;
    const long cell = face_cell(g, src, a, b);
    double* o = buf + (long)nv * e;
    for (int v = threadIdx.x; v < nv; v += blockDim.x)
      o[v] = (v < 5) ? g.w[v][cell] : g.w[5][cell * g.nchem + (v - 5)];
  }
}

// Ghost layers of face f in the reference's receive-buffer layout, from the descriptor
// the RHS kernel itself uses (so tests of this buffer test the kernel's ghost semantics).
__global__ void ghost_face_kernel(const FaceGeom g, const GhostFace G, double* __restrict__ dst, long nent)
{
  const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (e >= nent) return;
  const int nv = 5 + g.nchem;
  double* o = dst + (long)nv * e;
  if (G.mode == GHOST_BUF) {
    for (int v = 0; v < nv; v++) o[v] = G.buf[(long)nv * e + v];
    return;
  }
  int d; long a, b, na;
  face_decode(g, e, d, a, b, na);
  const long n = (g.f / 2 == 0) ? g.nx : (g.f / 2 == 1 ? g.ny : g.nz);
  const long pos = (g.f % 2 == 0) ? (long)d - 3 : n + d;
  const long cell = face_cell(g, G.a + (long)G.b * pos, a, b);
  for (int v = 0; v < 5; v++) {
    const double x = g.w[v][cell];
    o[v] = ((G.neg >> v) & 1u) ? -x : x;
  }
  for (int v = 0; v < g.nchem; v++) {
    const double x = g.w[5][cell * g.nchem + v];
    o[5 + v] = ((G.neg >> 5) & 1u) ? -x : x;
  }
}

#endif  // __CUDACC__ || EB_CUDA_EMU

}  // namespace eb
