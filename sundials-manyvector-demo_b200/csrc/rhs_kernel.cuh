// ---------------------------------------------------------------------------
// rhs_kernel.cuh -- the fused fluid right-hand side kernel (sm_100a).
//
// One launch evaluates  wdot = G - div F(w)  for every owned cell: what the reference
// does in five passes over memory (zero wdot, forcing, interior faces, boundary faces,
// divergence: /root/reference/src/utilities.cpp:28,65,76-116,123-195,198-245) with three
// flux scratch arrays (euler3D.hpp:497-499) is done here in a single pass that reads the
// state once and writes wdot once; face fluxes only ever exist in shared memory.
//
// Work decomposition (FP64-pipe bound, so the aim is: no redundant faces, no spills
// of the carried state, everything else on the other pipes):
//   * a CTA owns a (TX-1) x (TY-1) column of cells and marches along z over one
//     z-segment; thread (tx,ty) computes the LOWER x- and y-face of its cell and the
//     z-face above it.  The upper x/y faces come from the neighbouring threads through
//     shared memory; the last thread column/row of the tile only supplies those faces.
//   * the z-face below a cell is the one the same thread computed in the previous
//     step (kept in a thread-private shared-memory slot), so along z nothing is
//     computed twice except one face per segment.
//   * stencil values are read straight from global memory through L1 (each value is
//     reused by 6 faces x 3 directions; the FP64 pipe, not the LSU, is the limiter).
//   * ghost cells are never materialised for physical boundaries: each of the six
//     faces carries either an index map into the owned cells plus a sign mask
//     (periodic wrap / mirror / copy, euler3D.hpp:797-1166) or a pointer to the halo
//     buffer a neighbouring rank filled (reference wire layout, euler3D.hpp:648,696,744).
//
// Compiles under nvcc for the product and under g++ with tests/emu/cuda_emu.h for the
// CPU-side logic tests (never shipped).
// ---------------------------------------------------------------------------
#pragma once
#include "euler_math.cuh"

namespace eb {

enum { GHOST_MAP = 0, GHOST_BUF = 1 };

// Which fields a launch evaluates.  PART_ALL: one fused launch.  PART_FLUID / PART_TRACERS: the five
// fluid fields and the advected species in separate launches -- the species need only the face-local
// alpha and the normal velocities of the fluid part (utilities.cpp:368-380), which cost ~25 FP64
// instructions per face to rebuild from the per-cell arrays, and without the fluid face's ~170
// registers the species launch runs at a higher occupancy.
enum { PART_ALL = 0, PART_FLUID = 1, PART_TRACERS = 2 };

// How stencil positions beyond one side of the owned range along an axis are resolved.
struct GhostFace {
  int mode;            // GHOST_MAP: owned index = a + b*pos ; GHOST_BUF: halo buffer
  int b;
  long a;
  unsigned neg;        // bit v (0..4 fluid field, 5 = all tracers): negate the value
  const double* buf;   // GHOST_BUF: v + nv*(d + 3*(ta + na*tb)), d = layer 0..2
};

struct RhsParams {
  long nx, ny, nz;          // owned extents (EulerData::nxl,nyl,nzl)
  int nchem;
  int seg_len;              // cells per z-segment
  double gamma;
  double rdx, rdy, rdz;     // 1/dx, 1/dy, 1/dz
  double forcing[5];        // constant forcing assigned into wdot (external_forces hook)
  const double* w[6];       // rho, mx, my, mz, et (SoA), chem (AoS, species fastest)
  double* wdot[6];
  GhostFace ghost[6];       // W,E,S,N,B,F
  const double* aux[4];     // per-cell 1/rho, p, c, sqrt(rho) from aux_kernel (or all NULL:
                            // everything derived on the fly)
  int* state_flag;          // OR of legal_state failure bits (euler3D.hpp:1405-1414)
  // "slow RHS" mode of the multirate / IMEX drivers (fslow, multirate_chem_hydro_main.cpp:
  // 996-1083; fexpl, imex_chem_hydro_main.cpp:910-1000): before the fluxes the total energy is
  // rebuilt from the gas energy carried as the LAST chemistry species,
  //   et = chem[nchem-1]/EnergyUnits + |m|^2/(2 rho)            (:1033-1042, written into w)
  // and afterwards  chemdot[nchem-1] = etdot,  etdot = 0        (:1059-1068).
  int pair_sync;                      // rows of the tile rendezvous pairwise (see rhs_fused_kernel)
  int slow_mode;
  double inv_energy_units;
  double* et_rw;            // w[4] again, writable, for the rebuild (slow mode only)
  // sub-box of cells to evaluate: [lo, hi) per axis (the whole box for a single launch;
  // interior / boundary shells when the halo exchange is overlapped)
  long lo[3], hi[3];
};

// One resolved stencil point: where to read it and which fields change sign.
struct StencilPt {
  long off;        // owned: cell index; halo buffer: index of field 0
  unsigned neg;
  int src;         // -1 owned, else face id of the halo buffer
};

// Resolve the six points pos = idx-3 .. idx+2 along `dir` for the face whose index along
// that axis is idx (cell coordinates i,j,k hold idx in component dir).
template <bool GEN>
EB_HD void resolve(const RhsParams& P, int dir, long i, long j, long k, StencilPt pt[6])
{
  const long stride = (dir == 0) ? 1 : (dir == 1 ? P.nx : P.nx * P.ny);
  const long idx = (dir == 0) ? i : (dir == 1 ? j : k);
  const long cell = i + P.nx * (j + P.ny * k);   // may lie one past the end along dir
#pragma unroll
  for (int l = 0; l < 6; l++) {
    const long pos = idx - 3 + l;
    pt[l].off = cell + (l - 3) * stride;
    pt[l].neg = 0u;
    pt[l].src = -1;
    if (GEN) {
      const long n = (dir == 0) ? P.nx : (dir == 1 ? P.ny : P.nz);
      if (pos < 0 || pos >= n) {
        const int f = 2 * dir + (pos >= n ? 1 : 0);
        const GhostFace& G = P.ghost[f];
        if (G.mode == GHOST_MAP) {
          const long mapped = G.a + (long)G.b * pos;
          pt[l].off = cell + (mapped - idx) * stride;
          pt[l].neg = G.neg;
        } else {
          const long d = (pos < 0) ? pos + 3 : pos - n;
          const long ta = (dir == 0) ? j : i;
          const long tb = (dir == 2) ? j : k;
          const long na = (dir == 0) ? P.ny : P.nx;
          pt[l].off = (long)(5 + P.nchem) * (d + 3 * (ta + na * tb));
          pt[l].src = f;
        }
      }
    }
  }
}

// wdot is written once and never read by this kernel: a streaming store (evict-first in L2)
// leaves the L2 to the state lines neighbouring CTAs re-read (-0.27 % at 512^3/NVAR=15,
// profiles/r1g_ab_l1_hints.txt).
EB_HD void st_out(double* p, double x)
{
#if defined(__CUDA_ARCH__)
  __stcs(p, x);
#else
  *p = x;
#endif
}

template <bool GEN>
EB_HD double load_fluid(const RhsParams& P, const StencilPt& pt, int field)
{
  if (GEN) {
    double x = (pt.src < 0) ? P.w[field][pt.off] : P.ghost[pt.src].buf[pt.off + field];
    return ((pt.neg >> field) & 1u) ? -x : x;
  }
  return P.w[field][pt.off];
}

// Normal velocity u = mn/rho and sound speed c of one stencil point, for launches that evaluate
// the species only (PART_TRACERS): from the per-cell arrays where they are valid (owned points, and
// ghost points that are an owned cell with at most its momenta negated), else from the five fluid
// values of the point (halo slabs, Dirichlet ghosts, or no per-cell arrays at all).
template <bool GEN>
EB_HD void point_uc(const RhsParams& P, const StencilPt& pt, int fn, int f1, int f2, double& u, double& c)
{
  const bool owned = !GEN || pt.src < 0;
  if (P.aux[0] != nullptr && owned && (!GEN || (pt.neg & 0x11u) == 0u)) {
    u = load_fluid<GEN>(P, pt, fn) * P.aux[0][pt.off];
    c = P.aux[2][pt.off];
  } else {
    const double mn = load_fluid<GEN>(P, pt, fn);
    const CellAux a = cell_aux(P.gamma, load_fluid<GEN>(P, pt, 0), mn, load_fluid<GEN>(P, pt, f1),
                               load_fluid<GEN>(P, pt, f2), load_fluid<GEN>(P, pt, 4));
    u = mn * a.rinv;
    c = a.c;
  }
}

// Face flux of the fields of PART for the face below cell (i,j,k) along `dir`.  Each flux is
// handed to emit(v, value) with v in the reference's field order (rho,mx,my,mz,et,
// tracers...).  Returns the legal_state bits (euler3D.hpp:1405-1414) of stencil point 3,
// i.e. of cell (i,j,k) itself (0 from a species-only launch: the fluid launch reports them).
template <bool GEN, bool AG, int PART, class Emit>
EB_HD int face_all(const RhsParams& P, int dir, long i, long j, long k, Emit emit)
{
  StencilPt pt[6];
  resolve<GEN>(P, dir, i, j, k, pt);

  // sweep-aligned momentum order after the reference's swap (utilities.cpp:283-285)
  const int fn = 1 + dir;
  const int f1 = (dir == 1) ? 1 : 2;
  const int f2 = (dir == 2) ? 1 : 3;

  double alpha, u[6];
  int bits = 0;
  if (PART != PART_TRACERS) {
    FluidStencil s;
#pragma unroll
    for (int l = 0; l < 6; l++) {
      s.r[l] = load_fluid<GEN>(P, pt[l], 0);
      s.mn[l] = load_fluid<GEN>(P, pt[l], fn);
      s.m1[l] = load_fluid<GEN>(P, pt[l], f1);
      s.m2[l] = load_fluid<GEN>(P, pt[l], f2);
      s.e[l] = load_fluid<GEN>(P, pt[l], 4);
    }
    // Per-cell derived values: interior CTAs read what aux_kernel stored; CTAs that touch
    // ghost or halo points derive them from the (sign-mapped) state they just loaded.
    if (!GEN && P.aux[0] != nullptr) {
#pragma unroll
      for (int l = 0; l < 6; l++) {
        s.rinv[l] = P.aux[0][pt[l].off];
        s.p[l] = P.aux[1][pt[l].off];
        s.c[l] = P.aux[2][pt[l].off];
      }
      s.srL = P.aux[3][pt[2].off];
      s.srR = P.aux[3][pt[3].off];
    } else {
      // Tiles that touch a boundary (AG instantiation, chosen for launches where such tiles are
      // many: thin grids, small grids, the boundary shells of a decomposed run): owned points, and
      // ghost points that are an owned cell with at most its momenta negated (periodic wrap, Neumann,
      // reflecting: 1/rho, p, c, sqrt(rho) are even in the momenta), still read the per-cell arrays;
      // only Dirichlet ghosts (rho, e_t negated) and halo-slab points are derived here.
      const bool have_aux = AG && P.aux[0] != nullptr;
#pragma unroll
      for (int l = 0; l < 6; l++) {
        if (have_aux && pt[l].src < 0 && (pt[l].neg & 0x11u) == 0u) {
          s.rinv[l] = P.aux[0][pt[l].off];
          s.p[l] = P.aux[1][pt[l].off];
          s.c[l] = P.aux[2][pt[l].off];
          if (l == 2) s.srL = P.aux[3][pt[l].off];
          if (l == 3) s.srR = P.aux[3][pt[l].off];
        } else {
          const CellAux a = cell_aux(P.gamma, s.r[l], s.mn[l], s.m1[l], s.m2[l], s.e[l]);
          s.rinv[l] = a.rinv; s.p[l] = a.p; s.c[l] = a.c;
          if (l == 2) s.srL = a.sr;
          if (l == 3) s.srR = a.sr;
        }
      }
    }

    double f[5];
    fluid_face(s, P.gamma, f, alpha, u);
    bits = ((s.r[3] > 0.0) ? 0 : 1) | ((s.e[3] > 0.0) ? 0 : 2) | ((s.p[3] > 0.0) ? 0 : 4);

    emit(0, f[0]);
    emit(fn, f[1]);
    emit(f1, f[2]);
    emit(f2, f[3]);
    emit(4, f[4]);
  } else {
    // species-only launch: the face-local alpha = max_j |u_j| + c_j (utilities.cpp:368-380) and the
    // normal velocities, exactly as fluid_face forms them
    alpha = 0.0;
#pragma unroll
    for (int l = 0; l < 6; l++) {
      double c;
      point_uc<GEN>(P, pt[l], fn, f1, f2, u[l], c);
      const double a = fabs(u[l]) + c;
      alpha = (alpha < a) ? a : alpha;
    }
  }

  if (PART != PART_FLUID && P.nchem > 0) {
    double up[6], um[6];
#pragma unroll
    for (int l = 0; l < 6; l++) { up[l] = u[l] + alpha; um[l] = u[l] - alpha; }
    // per-point base pointer (and sign) of the tracer block of the cell
    const double* cp[6];
    double sg[6];
#pragma unroll
    for (int l = 0; l < 6; l++) {
      if (GEN && pt[l].src >= 0) cp[l] = P.ghost[pt[l].src].buf + pt[l].off + 5;
      else cp[l] = P.w[5] + pt[l].off * P.nchem;
      sg[l] = (GEN && ((pt[l].neg >> 5) & 1u)) ? -1.0 : 1.0;
    }
    // Tracers two at a time: one 16-byte load per stencil point feeds two independent
    // reconstructions (four WENO chains in flight), and the next pair is loaded while the
    // current one is reconstructed.  Needs nchem even and 16-byte aligned blocks, which
    // holds for the owned array (species-fastest, base from cudaMalloc) whenever nchem is
    // even; halo buffers hold NVAR = 5+nchem values per cell, so their tracer block is
    // only 8-byte aligned and takes the scalar path.
    bool vec = ((P.nchem & 1) == 0) && ((((unsigned long long)P.w[5]) & 15ull) == 0ull);
    if (GEN) {
#pragma unroll
      for (int l = 0; l < 6; l++) vec = vec && (pt[l].src < 0);
    }
    if (vec) {
      double2 c[6], cn[6];
#pragma unroll
      for (int l = 0; l < 6; l++) c[l] = *reinterpret_cast<const double2*>(cp[l]);
#pragma unroll 1
      for (int v = 0; v < P.nchem; v += 2) {
        const int vn = (v + 2 < P.nchem) ? v + 2 : v;
#pragma unroll
        for (int l = 0; l < 6; l++) cn[l] = *reinterpret_cast<const double2*>(cp[l] + vn);
        double a[6], b[6];
#pragma unroll
        for (int l = 0; l < 6; l++) {
          a[l] = GEN ? c[l].x * sg[l] : c[l].x;
          b[l] = GEN ? c[l].y * sg[l] : c[l].y;
        }
        const double fa = tracer_face(a, up, um);
        const double fb = tracer_face(b, up, um);
        emit(5 + v, fa);
        emit(6 + v, fb);
#pragma unroll
        for (int l = 0; l < 6; l++) c[l] = cn[l];
      }
    } else {
#pragma unroll 1
      for (int v = 0; v < P.nchem; v++) {
        double c[6];
#pragma unroll
        for (int l = 0; l < 6; l++) c[l] = GEN ? cp[l][v] * sg[l] : cp[l][v];
        emit(5 + v, tracer_face(c, up, um));
      }
    }
  }
  return bits;
}

template <bool AG, int PART, class Emit>
EB_HD int face_dispatch(bool gen, const RhsParams& P, int dir, long i, long j, long k, Emit emit)
{
  return gen ? face_all<true, AG, PART>(P, dir, i, j, k, emit) : face_all<false, AG, PART>(P, dir, i, j, k, emit);
}

#if defined(__CUDACC__) || defined(EB_CUDA_EMU)

// Pre-pass: per-cell 1/rho, p, c, sqrt(rho) for the cells [c0, c1) (40 B read, 32 B written
// per cell), so that the 18 stencils a cell sits in do not each redo a reciprocal and two
// square roots on the FP64 pipe.
__global__ void aux_kernel(const RhsParams P, double* a0, double* a1, double* a2, double* a3, long c0, long c1)
{
  for (long c = c0 + (long)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += (long)gridDim.x * blockDim.x) {
    const double r = P.w[0][c], mx = P.w[1][c], my = P.w[2][c], mz = P.w[3][c];
    double e = P.w[4][c];
    if (P.slow_mode) {
      e = P.w[5][c * P.nchem + (P.nchem - 1)] * P.inv_energy_units + 0.5 / r * (mx * mx + my * my + mz * mz);
      P.et_rw[c] = e;
    }
    if (a0 != nullptr) {
      const CellAux a = cell_aux(P.gamma, r, mx, my, mz, e);
      a0[c] = a.rinv; a1[c] = a.p; a2[c] = a.c; a3[c] = a.sr;
    }
  }
}

#if defined(__CUDACC__)
// named barrier: `count` threads (whole warps) meet at hardware barrier `id` (1..15)
__device__ __forceinline__ void eb_bar_sync(int id, int count)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif

// Dynamic shared memory: FX [NF][T-TX], FY [NF][T] (exchanged with the +x / +y neighbour
// thread) and ZLO [NF][T-TX] (thread-private: flux through the z-face below), NF the number of
// fields of the launch (NVAR, 5, or nchem); 127.5 KB at NF = 15 with 384 threads, which leaves
// the SM its 132 KB carve-out and 124 KB of L1 -- the stencil loads live in L1, and its size shows
// directly in the kernel time.
//
// Synchronisation.  Default: two CTA-wide barriers per plane (fluxes published / consumed).
// With P.pair_sync (tile rows are warps, TX == 32): FX never leaves the warp (lane l reads lane
// l+1's slot) and FY of row ty is read only by row ty-1, so each row meets just its two
// neighbours once per plane on named barriers (id ty with the row below, id ty+1 with the row
// above, 64 threads each) and FY is double-buffered (four arrays): row ty+1 can only overwrite a
// buffer after the NEXT rendezvous, which row ty reaches after it has consumed the buffer
// (pair_sync == 2: one FY buffer and a second rendezvous per plane instead, when the fourth
// array does not fit).
// Rows then drift apart by up to a phase per hop instead of all waiting for the slowest twice per
// plane, so their load bursts and FP64 stretches overlap.
// GW: the forcing is not the per-field constant of the config; the caller has run the
// external_forces hook into wdot (utilities.cpp:65) and every store is wdot = wdot - div.
// AG: boundary tiles read the per-cell arrays wherever they are valid (see face_all).
// PART: all fields, or the fluid fields / the species only (two launches, see the enum).
template <int MAXT, int MINB, bool GW = false, bool AG = false, int PART = PART_ALL>
__global__ void __launch_bounds__(MAXT, MINB) rhs_fused_kernel(const RhsParams P)
{
  EB_DYN_SMEM(double, smem);
  const int TX = blockDim.x, TY = blockDim.y, T = TX * TY;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * TX + tx;
  // fields of this launch: v0 .. v0+nf-1 in the reference's order
  const int v0 = (PART == PART_TRACERS) ? 5 : 0;
  const int nf = (PART == PART_ALL) ? 5 + P.nchem : (PART == PART_FLUID ? 5 : P.nchem);
  const bool pair = P.pair_sync != 0;
  const bool two_fy = P.pair_sync == 1;
  // the face-only top row of the tile never touches FX / ZLO: those arrays are [NF][TR]
  const int TR = T - TX;
  double* FX = smem + t - (long)v0 * TR;               // indexed with the global field number v
  double* FY = smem + (long)nf * TR + t - (long)v0 * T;
  long fy_flip = two_fy ? (long)nf * T : 0;         // signed distance to the other FY buffer
  double* ZLO = smem + (long)nf * (TR + (two_fy ? 2L : 1L) * T) + t - (long)v0 * TR;

  const long ti0 = P.lo[0] + (long)blockIdx.x * (TX - 1);
  const long tj0 = P.lo[1] + (long)blockIdx.y * (TY - 1);
  const long i = ti0 + tx, j = tj0 + ty;
  const long k0 = P.lo[2] + (long)blockIdx.z * P.seg_len;
  const long k1 = (k0 + P.seg_len < P.hi[2]) ? k0 + P.seg_len : P.hi[2];

  const bool row_ok = (ty < TY - 1) && (j < P.hi[1]);
  const bool col_ok = (tx < TX - 1) && (i < P.hi[0]);
  const bool owns = row_ok && col_ok;                 // this thread owns a cell column
  const bool need_x = row_ok && (i <= P.hi[0]);       // lower x-face at position i
  const bool need_y = col_ok && (j <= P.hi[1]);       // lower y-face at position j

  // CTA-uniform: does any stencil of this tile reach beyond the owned range in x / y?
  const bool gen_x = (ti0 - 3 < 0) || (ti0 + TX - 1 + 2 >= P.nx);
  const bool gen_y = (tj0 - 3 < 0) || (tj0 + TY - 1 + 2 >= P.ny);

  int mask = 0;

  // z-face below the first plane of the segment
  if (owns)
    face_dispatch<AG, PART>(k0 - 3 < 0 || k0 + 2 >= P.nz, P, 2, i, j, k0,
                  [&](int v, double x) { ZLO[v * TR] = x; });

  for (long k = k0; k < k1; k++) {
    // ---- phase A: lower x- and y-faces of plane k -> shared memory ----
    if (need_x) {
      const int bits = face_dispatch<AG, PART>(gen_x, P, 0, i, j, k, [&](int v, double x) { FX[v * TR] = x; });
      if (owns) mask |= bits;
    }
    if (need_y)
      face_dispatch<AG, PART>(gen_y, P, 1, i, j, k, [&](int v, double x) { FY[v * T] = x; });
    if (pair) {
      if (ty > 0) eb_bar_sync(ty, 64);
      if (ty < TY - 1) eb_bar_sync(ty + 1, 64);
    } else {
      __syncthreads();
    }

    // ---- phase B: z-face above cell (i,j,k); each flux closes the divergence of its
    //      field as soon as it exists (sum order of utilities.cpp:202-207) ----
    if (owns) {
      const long cell = i + P.nx * (j + P.ny * k);
      face_dispatch<AG, PART>(k + 1 - 3 < 0 || k + 1 + 2 >= P.nz, P, 2, i, j, k + 1,
                    [&](int v, double zup) {
                      const double div = ((FX[v * TR + 1] - FX[v * TR]) * P.rdx
                                        + (FY[v * T + TX] - FY[v * T]) * P.rdy)
                                        + (zup - ZLO[v * TR]) * P.rdz;
                      ZLO[v * TR] = zup;
                      if (GW) {
                        // the caller ran external_forces itself: wdot already holds G (utilities.cpp:65)
                        double* dst = (v < 5) ? P.wdot[v] + cell : P.wdot[5] + cell * P.nchem + (v - 5);
                        const double G = *dst;
                        if (!P.slow_mode) {
                          st_out(dst, G - div);
                        } else if (v == 4) {
                          st_out(P.wdot[5] + cell * P.nchem + (P.nchem - 1), G - div);
                          st_out(dst, 0.0);
                        } else if (v < 4 + P.nchem) {
                          st_out(dst, G - div);
                        }
                      } else if (!P.slow_mode) {
                        if (v < 5) st_out(P.wdot[v] + cell, P.forcing[v] - div);
                        else st_out(P.wdot[5] + cell * P.nchem + (v - 5), 0.0 - div);
                      } else if (v == 4) {          // etdot goes to the gas-energy species
                        st_out(P.wdot[5] + cell * P.nchem + (P.nchem - 1), P.forcing[4] - div);
                        st_out(P.wdot[4] + cell, 0.0);
                      } else if (v < 4) {
                        st_out(P.wdot[v] + cell, P.forcing[v] - div);
                      } else if (v < 4 + P.nchem) {  // every species but the last (overwritten above)
                        st_out(P.wdot[5] + cell * P.nchem + (v - 5), 0.0 - div);
                      }
                    });
    }
    if (two_fy) {
      __syncwarp();                 // FX of this plane consumed before the warp overwrites it
      FY += fy_flip;
      fy_flip = -fy_flip;
    } else if (pair) {              // single FY buffer: a second rendezvous releases it
      if (ty > 0) eb_bar_sync(ty, 64);
      if (ty < TY - 1) eb_bar_sync(ty + 1, 64);
    } else {
      __syncthreads();
    }
  }

  if (PART != PART_TRACERS && mask) atomicOr(P.state_flag, mask);
}

#endif  // __CUDACC__ || EB_CUDA_EMU

}  // namespace eb
