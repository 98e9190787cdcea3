// ---------------------------------------------------------------------------
// tracer_kernel.cuh -- the advected species of the fluid RHS as a kernel of their own
// (split mode, EULERB200_SPLIT=1; the fused kernel of rhs_kernel.cuh then runs with
// skip_tracers and only produces the five fluid fields).
//
// Why.  In the fused kernel the tracer half of the arithmetic (60 of the 70 WENO
// reconstructions per cell at NVAR = 15) runs at the occupancy the fluid half dictates (168
// registers, 12 warps per SM) and fetches its 6 x 80-byte stencil blocks through an L1 that the
// z-stencil working set overflows (profiles/README.md: L1 is the contested resource, every
// variant that adds L1 requests loses).  A tracer needs very little from the fluid
// (utilities.cpp:376-377,388,395,431,439,473): the normal velocity u_j of its six stencil cells
// and the face-local alpha = max_j (|u_j| + c_j).  Both are per-cell quantities, so this kernel
// takes them from the per-cell arrays aux_kernel writes (1/rho, c) and never touches the
// characteristic machinery.
//
// Work decomposition: one CTA = one (TX-1) x (TY-1) tile of cell columns, one z-segment and one
// PAIR of species (blockIdx.x = tile * npairs + pair: the pairs of a tile are neighbours in launch
// order, so that the 80-byte species blocks and the velocities they all read are L2 hits).  Marching along z:
//   * the z-stencil of the thread's cell column lives in a thread-private shared-memory RING of
//     six (c_a, c_b, u_z, |u_z|+c) tuples, one new plane fetched per step -- every tracer value
//     is read from global memory once per direction sweep instead of six times (a register ring
//     cost 48 move instructions per step and 48 registers: first version, profiles/);
//   * the x/y-stencils of plane k come from a shared-memory copy of the plane (tile plus the
//     3/2-cell arms of the stencil cross), double-buffered so that plane k+1 is filled while
//     plane k is being used: one CTA-wide barrier per plane;
//   * face fluxes are exchanged with the +x / +y neighbour thread through small shared arrays
//     (two species), again double-buffered by plane parity.
//   * interior tiles take a pipelined path: the global loads of a step (own cell and one arm
//     cell of plane k+1, own column of plane k+3; offsets only advance by a plane) are issued
//     before the x/y-face arithmetic and stored to shared memory after it.
// ~100 KB of shared memory per 256-thread CTA (two CTAs per SM), 150 KB per 384-thread CTA.
//
// Arithmetic: eb::tracer_face of euler_math.cuh, i.e. bit-for-bit what the fused kernel does.
// Compiles under nvcc and under g++ with tests/emu/cuda_emu.h.
// ---------------------------------------------------------------------------
#pragma once
#include "rhs_kernel.cuh"

namespace eb {

// Cell (i,j,k) with at most ONE coordinate outside the owned range (by at most the ghost depth):
// where to read it (mirrors resolve() for a single point).
EB_HD StencilPt resolve_cell(const RhsParams& P, long i, long j, long k, bool& valid)
{
  StencilPt pt;
  pt.off = i + P.nx * (j + P.ny * k);
  pt.neg = 0u;
  pt.src = -1;
  const bool ox = (i < 0 || i >= P.nx), oy = (j < 0 || j >= P.ny), oz = (k < 0 || k >= P.nz);
  valid = ((int)ox + (int)oy + (int)oz) <= 1;
  if (!valid || !(ox || oy || oz)) return pt;
  const int dir = ox ? 0 : (oy ? 1 : 2);
  const long n = (dir == 0) ? P.nx : (dir == 1 ? P.ny : P.nz);
  const long pos = (dir == 0) ? i : (dir == 1 ? j : k);
  const long stride = (dir == 0) ? 1 : (dir == 1 ? P.nx : P.nx * P.ny);
  if (pos < -3 || pos > n + 2) { valid = false; return pt; }
  const int f = 2 * dir + (pos >= n ? 1 : 0);
  const GhostFace& G = P.ghost[f];
  if (G.mode == GHOST_MAP) {
    const long mapped = G.a + (long)G.b * pos;
    pt.off += (mapped - pos) * stride;
    pt.neg = G.neg;
  } else {
    const long d = (pos < 0) ? pos + 3 : pos - n;
    const long ta = (dir == 0) ? j : i;
    const long tb = (dir == 2) ? j : k;
    const long na = (dir == 0) ? P.ny : P.nx;
    pt.off = (long)(5 + P.nchem) * (d + 3 * (ta + na * tb));
    pt.src = f;
  }
  return pt;
}

// What a tracer stencil needs from one cell: the two species values and, per requested
// direction d (bit d of DM), the velocity u_d = m_d / rho and s_d = |u_d| + c.
struct TracerCell { double ca, cb, u[3], s[3]; };

// GEN = false: the caller knows the cell is owned (CTA-uniform test), no ghost logic at all.
template <bool GEN, int DM>
EB_HD TracerCell fetch_tracer_cell(const RhsParams& P, long i, long j, long k, int v0, bool two)
{
  TracerCell C;
  StencilPt pt;
  if (GEN) {
    bool valid;
    pt = resolve_cell(P, i, j, k, valid);
    if (!valid) {                                 // outside the stencil cross: never used
      C.ca = C.cb = 0.0;
      for (int d = 0; d < 3; d++) C.u[d] = C.s[d] = 0.0;
      return C;
    }
  } else {
    pt.off = i + P.nx * (j + P.ny * k);
    pt.neg = 0u;
    pt.src = -1;
  }
  double rinv, c, m[3];
  if ((!GEN || (pt.src < 0 && pt.neg == 0u)) && P.aux[0] != nullptr) {
    rinv = P.aux[0][pt.off];
    c = P.aux[2][pt.off];
#pragma unroll
    for (int d = 0; d < 3; d++) m[d] = ((DM >> d) & 1) ? P.w[1 + d][pt.off] : 0.0;
  } else {                                        // ghost / halo cell: derive from the sign-mapped state
    const double r = load_fluid<GEN>(P, pt, 0);
    m[0] = load_fluid<GEN>(P, pt, 1);
    m[1] = load_fluid<GEN>(P, pt, 2);
    m[2] = load_fluid<GEN>(P, pt, 3);
    const double e = load_fluid<GEN>(P, pt, 4);
    const CellAux a = cell_aux(P.gamma, r, m[0], m[1], m[2], e);
    rinv = a.rinv;
    c = a.c;
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if ((DM >> d) & 1) {
      C.u[d] = m[d] * rinv;
      C.s[d] = fabs(C.u[d]) + c;
    } else {
      C.u[d] = C.s[d] = 0.0;
    }
  }
  const double* cp = (GEN && pt.src >= 0) ? P.ghost[pt.src].buf + pt.off + 5 : P.w[5] + pt.off * P.nchem;
  C.ca = cp[v0];
  C.cb = two ? cp[v0 + 1] : C.ca;
  if (GEN && ((pt.neg >> 5) & 1u) != 0u) { C.ca = -C.ca; C.cb = -C.cb; }
  return C;
}

// Flux of one species through one face from its six stencil values, the six velocities and
// s = |u| + c values (alpha as in fluid_face(): running maximum starting from 0).
EB_HD double tracer_flux6(const double c[6], const double u[6], const double s[6])
{
  double alpha = 0.0;
#pragma unroll
  for (int l = 0; l < 6; l++) alpha = (alpha < s[l]) ? s[l] : alpha;
  double up[6], um[6];
#pragma unroll
  for (int l = 0; l < 6; l++) { up[l] = u[l] + alpha; um[l] = u[l] - alpha; }
  return tracer_face(c, up, um);
}

// Shared-memory footprint in doubles: two plane copies (CA, CB on the stencil cross
// [TY+5][TX+5]; UX, SX [TY][TX+5]; UY, SY [TY+5][TX]), FX, FY [2 parities][2 species][T] and the
// z-ring [4 quantities][6 planes][T].
EB_HD long tracer_smem_doubles(int TX, int TY)
{
  const long WX = TX + 5, WY = TY + 5;
  return 2 * (2 * WX * WY + 2 * WX * TY + 2 * WY * TX) + (8L + 24L) * TX * TY;
}

struct TileGeom { int TX, TY, tx, ty, v0; bool two; long ti0, tj0; };

// Fill one shared plane copy with plane k: every thread its own cell; the cells of the stencil
// arms (3 below / 2 above the tile in x and y) go to the threads of the first five columns
// (x-arms of their row) and of the first five rows (y-arms of their column), or, for tiles
// thinner than that, to all threads in turn.
template <bool GEN>
EB_HD void tracer_fill(const RhsParams& P, const TileGeom& g, double* pl, long k)
{
  const int TX = g.TX, TY = g.TY, tx = g.tx, ty = g.ty, WX = TX + 5, WY = TY + 5;
  double* CA = pl;
  double* CB = CA + (long)WX * WY;
  double* UX = CB + (long)WX * WY;
  double* SX = UX + (long)WX * TY;
  double* UY = SX + (long)WX * TY;
  double* SY = UY + (long)WY * TX;
  {
    const TracerCell C = fetch_tracer_cell<GEN, 3>(P, g.ti0 + tx, g.tj0 + ty, k, g.v0, g.two);
    CA[(ty + 3) * WX + tx + 3] = C.ca;
    CB[(ty + 3) * WX + tx + 3] = C.cb;
    UX[ty * WX + tx + 3] = C.u[0];
    SX[ty * WX + tx + 3] = C.s[0];
    UY[(ty + 3) * TX + tx] = C.u[1];
    SY[(ty + 3) * TX + tx] = C.s[1];
  }
  auto x_arm = [&](int r, int q) {               // q = 0..4 -> columns 0,1,2 and TX+3,TX+4 of row r
    const int hx = (q < 3) ? q : TX + q;
    const TracerCell C = fetch_tracer_cell<GEN, 1>(P, g.ti0 - 3 + hx, g.tj0 + r, k, g.v0, g.two);
    CA[(r + 3) * WX + hx] = C.ca;
    CB[(r + 3) * WX + hx] = C.cb;
    UX[r * WX + hx] = C.u[0];
    SX[r * WX + hx] = C.s[0];
  };
  auto y_arm = [&](int cidx, int q) {            // q = 0..4 -> rows 0,1,2 and TY+3,TY+4 of column cidx
    const int hy = (q < 3) ? q : TY + q;
    const TracerCell C = fetch_tracer_cell<GEN, 2>(P, g.ti0 + cidx, g.tj0 - 3 + hy, k, g.v0, g.two);
    CA[hy * WX + cidx + 3] = C.ca;
    CB[hy * WX + cidx + 3] = C.cb;
    UY[hy * TX + cidx] = C.u[1];
    SY[hy * TX + cidx] = C.s[1];
  };
  if (TX >= 5 && TY >= 5) {
    if (tx < 5) x_arm(ty, tx);
    if (ty < 5) y_arm(tx, ty);
  } else {
    const int T = TX * TY, t = ty * TX + tx, nhx = 5 * TY, nh = nhx + 5 * TX;
    for (int h = t; h < nh; h += T) {
      if (h < nhx) x_arm(h / 5, h % 5);
      else y_arm((h - nhx) % TX, (h - nhx) / TX);
    }
  }
}

#if defined(__CUDACC__) || defined(EB_CUDA_EMU)

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) tracer_kernel(const RhsParams P)
{
  EB_DYN_SMEM(double, smem);
  const int TX = blockDim.x, TY = blockDim.y, T = TX * TY;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * TX + tx;
  const int WX = TX + 5, WY = TY + 5;
  const int npair = (P.nchem + 1) / 2;
  const int v0 = 2 * (int)(blockIdx.x % npair);     // pairs of one tile are neighbours in launch order
  const long seg = blockIdx.z;
  const bool two = v0 + 1 < P.nchem;

  const int cross = WX * WY;                                    // CA, CB
  const int plane_sz = 2 * cross + 2 * WX * TY + 2 * WY * TX;   // + UX, SX, UY, SY
  double* FXb = smem + 2 * plane_sz;        // [parity][species][T]
  double* FYb = FXb + 4 * T;
  double* RING = FYb + 4 * T + t;           // [c_a, c_b, u_z, s_z][slot 0..5][T], thread-private column t

  const long ti0 = P.lo[0] + (long)(blockIdx.x / npair) * (TX - 1);
  const long tj0 = P.lo[1] + (long)blockIdx.y * (TY - 1);
  const long i = ti0 + tx, j = tj0 + ty;
  const long k0 = P.lo[2] + seg * P.seg_len;
  const long k1 = (k0 + P.seg_len < P.hi[2]) ? k0 + P.seg_len : P.hi[2];

  const bool row_ok = (ty < TY - 1) && (j < P.hi[1]);
  const bool col_ok = (tx < TX - 1) && (i < P.hi[0]);
  const bool owns = row_ok && col_ok;
  const bool need_x = row_ok && (i <= P.hi[0]);
  const bool need_y = col_ok && (j <= P.hi[1]);

  // CTA-uniform: the tile with its stencil arms stays inside the owned range in x and y, every
  // arm cell finds a thread, and the per-cell arrays exist -> pipelined path with plain loads.
  const bool gen_xy = (ti0 - 3 < 0) || (ti0 + TX - 1 + 2 >= P.nx) || (tj0 - 3 < 0) || (tj0 + TY - 1 + 2 >= P.ny);
  const int nhx = 5 * TY, nh = nhx + 5 * TX;
  const bool fast = !gen_xy && nh <= T && P.aux[0] != nullptr;
  TileGeom tg;
  tg.TX = TX; tg.TY = TY; tg.tx = tx; tg.ty = ty; tg.ti0 = ti0; tg.tj0 = tj0; tg.v0 = v0; tg.two = two;

  // fast path: offsets that only advance by one plane per step, and where the thread's own
  // cell / its arm cell go in a plane copy
  const long ps = P.nx * P.ny;
  long off_c = i + P.nx * (j + P.ny * (k0 + 1));        // own cell in plane k+1
  long off_a = 0;
  int arm = 0, arm_c = 0, arm_u = 0;                    // arm: 0 none, 1 x-arm, 2 y-arm
  const int cen_c = (ty + 3) * WX + tx + 3;
  const int cen_ux = 2 * cross + ty * WX + tx + 3;
  const int cen_uy = 2 * cross + 2 * WX * TY + (ty + 3) * TX + tx;
  if (fast && t < nh) {
    if (t < nhx) {
      const int r = t / 5, q = t % 5, hx = (q < 3) ? q : TX + q;
      arm = 1;
      off_a = (ti0 - 3 + hx) + P.nx * ((tj0 + r) + P.ny * (k0 + 1));
      arm_c = (r + 3) * WX + hx;
      arm_u = 2 * cross + r * WX + hx;
    } else {
      const int g = t - nhx, cidx = g % TX, q = g / TX, hy = (q < 3) ? q : TY + q;
      arm = 2;
      off_a = (ti0 + cidx) + P.nx * ((tj0 - 3 + hy) + P.ny * (k0 + 1));
      arm_c = hy * WX + cidx + 3;
      arm_u = 2 * cross + 2 * WX * TY + hy * TX + cidx;
    }
  }
  const int arm_s = arm_u + (arm == 1 ? WX * TY : WY * TX);
  // loop-invariant shared-memory indices of the x/y stencils and of the ring arrays
  const int bx = (ty + 3) * WX + tx, bux = ty * WX + tx, by = ty * WX + tx + 3, buy = ty * TX + tx;
  const int T6 = 6 * T;
  const double* arm_m = (arm == 1) ? P.w[1] : P.w[2];

  // z-ring of the thread's cell column in shared memory: slot (head + l) % 6 holds stencil
  // point l of the z-face about to be evaluated
  int head = 0;
  double zlo_a = 0.0, zlo_b = 0.0;
  if (owns) {
    double za[6], zb[6], zu[6], zs[6];
#pragma unroll
    for (int l = 0; l < 6; l++) {
      const TracerCell C = fetch_tracer_cell<true, 4>(P, i, j, k0 - 3 + l, v0, two);
      za[l] = C.ca; zb[l] = C.cb; zu[l] = C.u[2]; zs[l] = C.s[2];
      RING[(0 * 6 + l) * T] = C.ca;
      RING[(1 * 6 + l) * T] = C.cb;
      RING[(2 * 6 + l) * T] = C.u[2];
      RING[(3 * 6 + l) * T] = C.s[2];
    }
    zlo_a = tracer_flux6(za, zu, zs);
    zlo_b = tracer_flux6(zb, zu, zs);
  }
  if (gen_xy) tracer_fill<true>(P, tg, smem, k0); else tracer_fill<false>(P, tg, smem, k0);
  __syncthreads();

  for (long k = k0; k < k1; k++) {
    const int par = (int)((k - k0) & 1);
    const double* pl = smem + par * plane_sz;
    const double* CA = pl;
    const double* CB = CA + cross;
    const double* UX = CB + cross;
    const double* SX = UX + WX * TY;
    const double* UY = SX + WX * TY;
    const double* SY = UY + WY * TX;
    double* FX = FXb + par * 2 * T + t;
    double* FY = FYb + par * 2 * T + t;
    double* nxt = smem + (1 - par) * plane_sz;
    const bool have_next = k + 1 < k1;

    // ---- loads of this step, issued before the arithmetic that hides their latency ----
    double Lca = 0, Lcb = 0, Lmx = 0, Lmy = 0, Lri = 0, Lc = 0;      // own cell, plane k+1
    double Aca = 0, Acb = 0, Am = 0, Ari = 0, Ac = 0;                // arm cell, plane k+1
    double Rca = 0, Rcb = 0, Ru = 0, Rs = 0, Rri = 0, Rc = 0;        // ring: own column, plane k+3
    const bool ring_plain = k + 3 < P.nz;
    if (fast && have_next) {
      const double* cp = P.w[5] + off_c * P.nchem + v0;
      Lca = cp[0];
      Lcb = two ? cp[1] : Lca;
      Lmx = P.w[1][off_c]; Lmy = P.w[2][off_c]; Lri = P.aux[0][off_c]; Lc = P.aux[2][off_c];
      if (arm) {
        const double* ap = P.w[5] + off_a * P.nchem + v0;
        Aca = ap[0];
        Acb = two ? ap[1] : Aca;
        Am = arm_m[off_a]; Ari = P.aux[0][off_a]; Ac = P.aux[2][off_a];
      }
    }
    if (owns) {
      if (fast && ring_plain) {
        const long off_r = off_c + 2 * ps;
        const double* rp = P.w[5] + off_r * P.nchem + v0;
        Rca = rp[0];
        Rcb = two ? rp[1] : Rca;
        Ru = P.w[3][off_r]; Rri = P.aux[0][off_r]; Rc = P.aux[2][off_r];
      } else {
        const TracerCell C = ring_plain ? fetch_tracer_cell<false, 4>(P, i, j, k + 3, v0, two)
                                        : fetch_tracer_cell<true, 4>(P, i, j, k + 3, v0, two);
        Rca = C.ca; Rcb = C.cb; Ru = C.u[2]; Rs = C.s[2];
      }
    }

    // ---- phase A: lower x- and y-face of plane k from the shared plane copy ----
    if (need_x) {
      double c[6], u[6], s[6];
#pragma unroll
      for (int l = 0; l < 6; l++) {
        c[l] = CA[bx + l];
        u[l] = UX[bux + l];
        s[l] = SX[bux + l];
      }
      FX[0] = tracer_flux6(c, u, s);
#pragma unroll
      for (int l = 0; l < 6; l++) c[l] = CB[bx + l];
      FX[T] = tracer_flux6(c, u, s);
    }
    if (need_y) {
      double c[6], u[6], s[6];
#pragma unroll
      for (int l = 0; l < 6; l++) {
        c[l] = CA[by + l * WX];
        u[l] = UY[buy + l * TX];
        s[l] = SY[buy + l * TX];
      }
      FY[0] = tracer_flux6(c, u, s);
#pragma unroll
      for (int l = 0; l < 6; l++) c[l] = CB[by + l * WX];
      FY[T] = tracer_flux6(c, u, s);
    }

    // ---- the loaded values go where the next step reads them ----
    if (have_next) {
      if (fast) {
        const double ux = Lmx * Lri, uy = Lmy * Lri;
        nxt[cen_c] = Lca;
        nxt[cross + cen_c] = Lcb;
        nxt[cen_ux] = ux;
        nxt[cen_ux + WX * TY] = fabs(ux) + Lc;
        nxt[cen_uy] = uy;
        nxt[cen_uy + WY * TX] = fabs(uy) + Lc;
        if (arm) {
          const double ua = Am * Ari;
          nxt[arm_c] = Aca;
          nxt[cross + arm_c] = Acb;
          nxt[arm_u] = ua;
          nxt[arm_s] = fabs(ua) + Ac;
        }
      } else if (gen_xy) {
        tracer_fill<true>(P, tg, nxt, k + 1);
      } else {
        tracer_fill<false>(P, tg, nxt, k + 1);
      }
    }
    if (owns) {
      if (fast && ring_plain) { Ru = Ru * Rri; Rs = fabs(Ru) + Rc; }
      double* slot = RING + head * T;          // overwrites the oldest plane
      slot[0] = Rca;
      slot[T6] = Rcb;
      slot[2 * T6] = Ru;
      slot[3 * T6] = Rs;
      head = (head == 5) ? 0 : head + 1;
    }
    off_c += ps;
    off_a += ps;
    __syncthreads();

    // ---- phase B: z-face above the cell, divergence, store ----
    if (owns) {
      double za[6], zb[6], zu[6], zs[6];
#pragma unroll
      for (int l = 0; l < 6; l++) {
        const double* slot = RING + ((head + l >= 6) ? head + l - 6 : head + l) * T;
        za[l] = slot[0];
        zb[l] = slot[T6];
        zu[l] = slot[2 * T6];
        zs[l] = slot[3 * T6];
      }
      const double zup_a = tracer_flux6(za, zu, zs);
      const double zup_b = tracer_flux6(zb, zu, zs);
      const long cell = i + P.nx * (j + P.ny * k);
      double* out = P.wdot[5] + cell * P.nchem + v0;
      const double div_a = ((FX[1] - FX[0]) * P.rdx + (FY[TX] - FY[0]) * P.rdy) + (zup_a - zlo_a) * P.rdz;
      st_out(out, 0.0 - div_a);
      if (two) {
        const double div_b = ((FX[T + 1] - FX[T]) * P.rdx + (FY[T + TX] - FY[T]) * P.rdy) + (zup_b - zlo_b) * P.rdz;
        st_out(out + 1, 0.0 - div_b);
      }
      zlo_a = zup_a;
      zlo_b = zup_b;
    }
  }
}

#endif  // __CUDACC__ || EB_CUDA_EMU

}  // namespace eb
