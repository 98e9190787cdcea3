// ---------------------------------------------------------------------------
// host_setup.h -- host-side arithmetic behind the C ABI: domain decomposition,
// ghost-face descriptors, launch geometry.  No CUDA calls in here (the CPU logic tests
// include it too).
// ---------------------------------------------------------------------------
#pragma once
#include <cmath>
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "../../include/eulerb200.h"
#include "rhs_kernel.cuh"

namespace eb {

// MPI_Dims_create as every mainstream MPI answers it: a balanced factorisation in
// non-increasing order (used by EulerData::SetupDecomp, euler3D.hpp:430).
inline void dims_create(int nnodes, int ndims, int* dims)
{
  std::vector<int> primes;
  int n = nnodes;
  for (int p = 2; p * p <= n; p++)
    while (n % p == 0) { primes.push_back(p); n /= p; }
  if (n > 1) primes.push_back(n);
  std::sort(primes.begin(), primes.end(), [](int a, int b) { return a > b; });
  std::vector<int> bins(ndims, 1);
  for (size_t q = 0; q < primes.size() && ndims > 0; q++) {
    int best = 0;
    for (int d = 1; d < ndims; d++)
      if (bins[d] < bins[best]) best = d;
    bins[best] *= primes[q];
  }
  std::sort(bins.begin(), bins.end(), [](int a, int b) { return a > b; });
  for (int d = 0; d < ndims; d++) dims[d] = bins[d];
}

// euler3D.hpp:416-440,443-457,466-494,505-567
inline int decompose(int nprocs, int rank, const int64_t* n, const int32_t* bc,
                     int32_t* dims, int32_t* coords, int64_t* ext, int32_t* nbr)
{
  for (int d = 0; d < 3; d++)
    if ((bc[2 * d] == EULERB200_BC_PERIODIC) != (bc[2 * d + 1] == EULERB200_BC_PERIODIC)) return 1;
  int truedims = 0;
  for (int d = 0; d < 3; d++) truedims += (n[d] > 3) ? 1 : 0;
  int sugg[3] = {1, 1, 1};
  dims_create(nprocs, truedims, sugg);
  int q = 0;
  for (int d = 0; d < 3; d++) dims[d] = (n[d] > 3) ? sugg[q++] : 1;
  if (dims[0] * dims[1] * dims[2] != nprocs) return -1;   // e.g. nprocs > 1 on a 3x3x3 grid
  // non-reordered Cartesian communicator: row-major ranks, coords[0] slowest
  int r = rank;
  coords[2] = r % dims[2]; r /= dims[2];
  coords[1] = r % dims[1]; r /= dims[1];
  coords[0] = r;
  for (int d = 0; d < 3; d++) {
    ext[2 * d] = n[d] * coords[d] / dims[d];
    ext[2 * d + 1] = n[d] * (coords[d] + 1) / dims[d] - 1;
    if (ext[2 * d + 1] - ext[2 * d] + 1 < 3) return -1;
  }
  for (int d = 0; d < 3; d++)
    for (int side = 0; side < 2; side++) {
      const bool periodic = bc[2 * d] == EULERB200_BC_PERIODIC;
      const bool inner = side == 0 ? coords[d] > 0 : coords[d] < dims[d] - 1;
      int v = EULERB200_NO_NEIGHBOR;
      if (inner || periodic) {
        int c[3] = {coords[0], coords[1], coords[2]};
        c[d] = (c[d] + (side == 0 ? -1 : 1) + dims[d]) % dims[d];
        v = (c[0] * dims[1] + c[1]) * dims[2] + c[2];
      }
      nbr[2 * d + side] = v;
    }
  return 0;
}

inline int64_t face_len(const eulerb200_config& c, int f)
{
  const int64_t nv = 5 + c.nchem;
  if (f < 2) return nv * 3 * c.nyl * c.nzl;
  if (f < 4) return nv * 3 * c.nxl * c.nzl;
  return nv * 3 * c.nxl * c.nyl;
}

inline bool face_is_remote(const eulerb200_config& c, int f)
{
  return c.nbr[f] != EULERB200_NO_NEIGHBOR && c.nbr[f] != c.rank;
}

// Ghost descriptor of face f.  Physical boundaries (euler3D.hpp:797-1166): low side
// mirrors (ghost -1-m <- own m), high side COPIES (ghost n+m <- own n-3+m); reflecting
// negates the face-normal momentum, Dirichlet everything.  A periodic wrap onto the same
// rank is an index shift by n.  Anything else reads the halo buffer `recv`.
inline int ghost_face(const eulerb200_config& c, int f, const double* recv, GhostFace* g)
{
  const int dir = f / 2, side = f % 2;
  const long n = dir == 0 ? c.nxl : (dir == 1 ? c.nyl : c.nzl);
  g->buf = nullptr;
  g->neg = 0u;
  if (c.nbr[f] == EULERB200_NO_NEIGHBOR) {
    g->mode = GHOST_MAP;
    if (side == 0) { g->a = -1; g->b = -1; } else { g->a = -3; g->b = 1; }
    switch (c.bc[f]) {
      case EULERB200_BC_NEUMANN: break;
      case EULERB200_BC_REFLECTING: g->neg = 1u << (1 + dir); break;
      case EULERB200_BC_DIRICHLET: g->neg = 0x3Fu; break;
      default: return -1;   // a periodic face always has a neighbour (possibly this rank)
    }
  } else if (c.nbr[f] == c.rank) {
    g->mode = GHOST_MAP;
    g->b = 1;
    g->a = side == 0 ? n : -n;
  } else {
    g->mode = GHOST_BUF;
    g->a = 0; g->b = 0;
    g->buf = recv;
  }
  return 0;
}

// Order of the point-to-point operations of one halo exchange.  Sends go out in face order
// W,E,S,N,B,F; the receive posted next to the send of face f is the one for the OPPOSITE
// face f^1.  Between any two ranks the k-th send of one then meets the k-th receive of the
// other, which is what the message tags do in the reference (euler3D.hpp:608-640,663-784).
struct ExchangeOp { int32_t kind, face, peer; };   // kind 0 = send, 1 = recv
inline int exchange_plan(const eulerb200_config& c, ExchangeOp* ops)
{
  int n = 0;
  for (int f = 0; f < 6; f++) {
    if (face_is_remote(c, f)) { ops[n].kind = 0; ops[n].face = f; ops[n].peer = c.nbr[f]; n++; }
    const int r = f ^ 1;
    if (face_is_remote(c, r)) { ops[n].kind = 1; ops[n].face = r; ops[n].peer = c.nbr[r]; n++; }
  }
  return n;
}

struct LaunchGeom {
  unsigned gx, gy, gz;
  int tx, ty;
  int seg_len;
  size_t smem;
  int pair;          // 0: CTA-wide barriers, 1: row rendezvous + two FY buffers, 2: row rendezvous twice per plane
  int xc;            // tiles own all tx columns; the closing x-face column comes from the top warp (rhs_fused_kernel<..., XC>)
};

// Tile shape: 256 threads; a full warp along x whenever the box is at least 31 cells
// wide, otherwise the narrowest power of two that holds extent+1 faces (thin boxes such
// as the 3-cell-wide hurricane plane then put the threads along y).
// nf: fields the launch evaluates (NVAR for the fused launch, 5 / nchem for the fluid / species launches)
// want_xc: tiles of tx owned columns (rhs_fused_kernel<..., XC>) where rows are warps
inline LaunchGeom launch_geom(const long lo[3], const long hi[3], int nf, int threads = 256, int want_pair = 0,
                              long ctas_target = 5920, int want_xc = 0)
{
  LaunchGeom L;
  const long ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  int tx = 32;
  while (tx > 2 && tx / 2 >= ex + 1) tx /= 2;
  int ty = threads / tx;
  while (ty > 2 && ty / 2 >= ey + 1) ty /= 2;
  // three flux arrays [NVAR][threads] must fit the 227 KB of shared memory a CTA can opt in to:
  // many species (NVAR up to 64) get a flatter tile
  // (FX and ZLO skip the face-only top row: 3 ty - 2 rows of NVAR x tx doubles in all)
  const size_t smem_max = (size_t)227 * 1024, per_row = (size_t)nf * tx * sizeof(double);
  while (ty > 2 && per_row * (3 * ty - 2) > smem_max) ty--;
  L.tx = tx; L.ty = ty;
  L.xc = (want_xc && tx == 32 && ty >= 2) ? 1 : 0;
  L.gx = L.xc ? (unsigned)((ex + tx - 1) / tx) : (unsigned)((ex + tx - 2) / (tx - 1));
  L.gy = (unsigned)((ey + ty - 2) / (ty - 1));
  // z-segments: enough CTAs for ~20 waves on 148 SMs x 2 resident CTAs (ctas_target; tuning knob
  // EULERB200_CTAS), but at least 8 cells per segment (one extra z-face is computed per segment)
  const long tiles = (long)L.gx * L.gy;
  long nseg = (ctas_target + tiles - 1) / tiles;
  nseg = std::max(1L, std::min(nseg, (ez + 7) / 8));
  {
    // around that count, the one that wastes least: the last wave of CTAs on the 148 SMs (one CTA per SM)
    // is only partly filled, and every segment computes one z-face more than it has planes
    // (512^3, 16 x 47 tiles: 12 segments = 60.97 waves instead of 8 = 40.65; measured 59.0 vs 59.2 ms)
    const double sms = 148.0;
    double best = 1e300;
    long pick = nseg;
    for (long g = std::max(1L, nseg / 2); g <= std::min(2 * nseg, (ez + 7) / 8); g++) {
      const long seg = (ez + g - 1) / g, gz = (ez + seg - 1) / seg;
      const double ideal = (double)(tiles * gz) / sms;
      const double cost = std::ceil(ideal) / ideal * (1.0 + 1.0 / (3.0 * (double)seg));
      if (cost < best - 1e-12) { best = cost; pick = g; }
    }
    nseg = pick;
  }
  L.seg_len = (int)((ez + nseg - 1) / nseg);
  L.gz = (unsigned)((ez + L.seg_len - 1) / L.seg_len);
  // pairwise row rendezvous (rhs_fused_kernel): rows must be warps, one named barrier per row
  // pair (ids 1..15: at most 16 rows), and the second FY buffer has to fit the 227 KB a CTA can have
  L.pair = (tx == 32 && ty >= 2 && ty <= 16) ? want_pair : 0;
  if (L.pair == 1 && per_row * (4 * ty - 2) > smem_max) L.pair = 2;
  L.smem = per_row * ((L.pair == 1 ? 4 : 3) * ty - 2);
  // XC: two more slots per FX row (the face column of the top warp, double-buffered) and the four mbarriers
  // of its hand-over
  if (L.xc) L.smem += (size_t)nf * 2 * (ty - 1) * sizeof(double) + 32;
  return L;
}

// Fraction of the tiles of a launch whose x- or y-stencils reach beyond the owned range (the
// CTA-uniform gen_x / gen_y test of rhs_fused_kernel).
inline double boundary_tile_fraction(const long lo[3], const long hi[3], long nx, long ny, const LaunchGeom& L)
{
  (void)hi;
  long okx = 0, oky = 0;
  const int px = L.xc ? L.tx : L.tx - 1;           // columns a tile owns
  for (unsigned b = 0; b < L.gx; b++) {
    const long t0 = lo[0] + (long)b * px;
    if (!(t0 - 3 < 0 || t0 + px + 2 >= nx)) okx++;
  }
  for (unsigned b = 0; b < L.gy; b++) {
    const long t0 = lo[1] + (long)b * (L.ty - 1);
    if (!(t0 - 3 < 0 || t0 + L.ty - 1 + 2 >= ny)) oky++;
  }
  return 1.0 - ((double)okx / L.gx) * ((double)oky / L.gy);
}

// The launches of one RHS of a rank with remote neighbours (rhs_impl): box 0 is the interior -- every cell
// whose stencils stay clear of the remote faces, evaluated while the halo is on its way -- boxes 1.. are
// the boundary shells (z-low, z-high, y-low, y-high, x-low, x-high; empty ones left out), evaluated once
// it is in.  Together they cover the box exactly once.
// A shell holds the three layers next to its face.  thick (EULERB200_SHELLS=1): in x and y it is one TILE thick
// instead (pitch[d] = columns / rows a tile owns) and cut where the tile grid of the whole box has a seam, so
// that interior and shell launches together run the tiles a single launch over the box would run -- full
// tiles, no 3-of-4-column CTAs; thin boxes (fewer than four tiles along an axis) keep three layers, and along z a
// shell is always three planes (z-segments are not a tiling).  Measured at 512^3 / NVAR 15 on 2 GPUs: the
// 32-column shell takes 4.9 ms after a 53.7 ms interior, 60.42 ms in all against 60.20 with the 3-column shell
// (57.4 + 0.93): the 2.5 x per-cell cost of a shell is in its boundary tiles (halo reads, short z-segments),
// not in the thin tiles -- so three layers stay the default.
struct BoxList { int count; long lo[7][3], hi[7][3]; };
inline bool overlap_boxes(const long n[3], const bool remote[6], const long pitch[3], bool thick, BoxList* B)
{
  long lo[3], hi[3];
  bool interior = true;
  for (int d = 0; d < 3; d++) {
    const bool tile = thick && d < 2 && pitch[d] >= 3 && n[d] >= 4 * pitch[d];
    lo[d] = remote[2 * d] ? (tile ? pitch[d] : 3) : 0;
    hi[d] = n[d];
    if (remote[2 * d + 1]) hi[d] = tile ? lo[d] + pitch[d] * ((n[d] - 3 - lo[d]) / pitch[d]) : n[d] - 3;
    if (hi[d] <= lo[d]) interior = false;
  }
  B->count = 0;
  if (!interior) return false;
  auto add = [&](long x0, long x1, long y0, long y1, long z0, long z1) {
    if (x1 <= x0 || y1 <= y0 || z1 <= z0) return;
    const int q = B->count++;
    B->lo[q][0] = x0; B->hi[q][0] = x1; B->lo[q][1] = y0; B->hi[q][1] = y1; B->lo[q][2] = z0; B->hi[q][2] = z1;
  };
  add(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);        // interior
  add(0, n[0], 0, n[1], 0, lo[2]);                      // z-low
  add(0, n[0], 0, n[1], hi[2], n[2]);                   // z-high
  add(0, n[0], 0, lo[1], lo[2], hi[2]);                 // y-low
  add(0, n[0], hi[1], n[1], lo[2], hi[2]);              // y-high
  add(0, lo[0], lo[1], hi[1], lo[2], hi[2]);            // x-low
  add(hi[0], n[0], lo[1], hi[1], lo[2], hi[2]);         // x-high
  return true;
}

}  // namespace eb
