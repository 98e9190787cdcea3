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
#ifdef EB_STRICT
#include "strict_face.cuh"
#endif

// EB_ABLATE: bit mask of measurement-only ablations used by tools/ablate.cu to attribute kernel time
// (1 no barriers, 2 species values synthesised instead of loaded, 4 results not stored, 8 the
// species launch's u and c synthesised instead of loaded).  Never defined in the product build.
#ifndef EB_ABLATE
#define EB_ABLATE 0
#endif

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
  double rdx, rdy, rdz;     // EB_RD_SCALE / dx, dy, dz: the inverse spacings, halved where the faces hand out twice the flux (EB_FOLD_HALF)
  double dx, dy, dz;        // (the strict build divides as the reference does, utilities.cpp:202-207)
  double forcing[5];        // constant forcing assigned into wdot (external_forces hook)
  const double* w[6];       // rho, mx, my, mz, et (SoA), chem (AoS, species fastest)
  double* wdot[6];
  GhostFace ghost[6];       // W,E,S,N,B,F
  const double* aux[4];     // per-cell 1/rho, p, c, sqrt(rho) from aux_kernel (or all NULL:
                            // everything derived on the fly)
  const double* chemT;      // pair-interleaved copy of the species made by aux_kernel (or NULL):
                            // species 2q, 2q+1 of cell c at chemT[2 (q N + c)], N = nx ny nz.  The
                            // vector the drivers own is species-fastest (euler3D.hpp:65): a warp
                            // reading one species pair of 32 neighbouring cells from it touches 20
                            // cache lines (stride 80 B at nchem = 10), from this copy 4-5.
  int* state_flag;          // OR of legal_state failure bits (euler3D.hpp:1405-1414)
  // "slow RHS" mode of the multirate / IMEX drivers (fslow, multirate_chem_hydro_main.cpp:
  // 996-1083; fexpl, imex_chem_hydro_main.cpp:910-1000): before the fluxes the total energy is
  // rebuilt from the gas energy carried as the LAST chemistry species,
  //   et = chem[nchem-1]/EnergyUnits + |m|^2/(2 rho)            (:1033-1042, written into w by aux_kernel)
  // and afterwards  chemdot[nchem-1] = etdot,  etdot = 0        (:1059-1068, slow_post_kernel).
  int pair_sync;                      // rows of the tile rendezvous pairwise (see rhs_fused_kernel)
  int vec_store;                      // species pairs go out as one 16-byte store (nchem even, wdot[5] 16-byte aligned)
  int slow_mode;
  double inv_energy_units;
  double* et_rw;            // w[4] again, writable, for the rebuild (slow mode only)
  // sub-box of cells to evaluate: [lo, hi) per axis (the whole box for a single launch;
  // interior / boundary shells when the halo exchange is overlapped)
  long lo[3], hi[3];
};

// One resolved stencil point: where to read it and which fields change sign.
struct StencilPt {
  unsigned off;    // owned: cell index; halo buffer: index of field 0.  32 bits on purpose: one
                   // IMAD.WIDE forms base + 8 off (eulerb200_create refuses boxes of 2^31 cells or more)
  unsigned neg;
  int src;         // -1 owned, else face id of the halo buffer
};

// Resolve the six points pos = idx-3 .. idx+2 along `dir` for the face whose index along
// that axis is idx (cell coordinates i,j,k hold idx in component dir).
// Index arithmetic in 32 bits throughout (one IMAD.WIDE then forms base + 8 off per load; in 64 bits
// every load costs an add-with-carry pair): eulerb200_create refuses boxes of 2^31 cells or more.
template <bool GEN>
EB_HD void resolve(const RhsParams& P, int dir, int i, int j, int k, StencilPt pt[6])
{
  const int nx = (int)P.nx, ny = (int)P.ny;
  const int stride = (dir == 0) ? 1 : (dir == 1 ? nx : nx * ny);
  const int idx = (dir == 0) ? i : (dir == 1 ? j : k);
  const int cell = i + nx * (j + ny * k);   // may lie one past the end along dir
#pragma unroll
  for (int l = 0; l < 6; l++) {
    const int pos = idx - 3 + l;
    pt[l].off = (unsigned)(cell + (l - 3) * stride);
    pt[l].neg = 0u;
    pt[l].src = -1;
    if (GEN) {
      const int n = (dir == 0) ? nx : (dir == 1 ? ny : (int)P.nz);
      if (pos < 0 || pos >= n) {
        const int f = 2 * dir + (pos >= n ? 1 : 0);
        const GhostFace& G = P.ghost[f];
        if (G.mode == GHOST_MAP) {
          const int mapped = (int)G.a + G.b * pos;
          pt[l].off = (unsigned)(cell + (mapped - idx) * stride);
          pt[l].neg = G.neg;
        } else {
          const int d = (pos < 0) ? pos + 3 : pos - n;
          const int ta = (dir == 0) ? j : i;
          const int tb = (dir == 2) ? j : k;
          const int na = (dir == 0) ? ny : nx;
          pt[l].off = (unsigned)((5 + P.nchem) * (d + 3 * (ta + na * tb)));
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
  if (EB_ABLATE & 4) { if (x == 1.234567e300) *p = x; return; }
  __stcs(p, x);
#else
  *p = x;
#endif
}

EB_HD void st_out2(double* p, double x, double y)      // p 16-byte aligned
{
#if defined(__CUDA_ARCH__)
  if (EB_ABLATE & 4) { if (x == 1.234567e300) *p = y; return; }
  __stcs(reinterpret_cast<double2*>(p), make_double2(x, y));
#else
  p[0] = x; p[1] = y;
#endif
}

template <bool GEN>
EB_HD double load_fluid(const RhsParams& P, const StencilPt& pt, int field)
{
  if (EB_ABLATE & 16) return (field == 0 || field == 4) ? 1.5 + 1e-9 * (double)pt.off : 0.1 + 1e-9 * (double)(pt.off + field);
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
  if (EB_ABLATE & 8) { u = 0.1 + 1e-9 * (double)pt.off; c = 1.0 + 1e-10 * (double)pt.off; return; }
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
// STG: the species values of the six stencil points come from the shared-memory copy of the plane
// (rhs_fused_kernel<..., STAGE>): pair q of point l at sb[l * ss + q].
template <bool GEN, bool AG, int PART, class Emit, bool STG = false>
EB_HD int face_all(const RhsParams& P, int dir, int i, int j, int k, Emit emit, const double2* sb = nullptr, int ss = 0)
{
  StencilPt pt[6];
  resolve<GEN>(P, dir, i, j, k, pt);

  // sweep-aligned momentum order after the reference's swap (utilities.cpp:283-285)
  const int fn = 1 + dir;
  const int f1 = (dir == 1) ? 1 : 2;
  const int f2 = (dir == 2) ? 1 : 3;

#ifdef EB_STRICT
  {
    // the reference's arithmetic, operation for operation (strict_face.cuh): gather the stencil as
    // pack1D_* does (euler3D.hpp:1197-1378), face_flux, hand the fluxes on
    const int nvar = 5 + P.nchem;
    double s[6][strict::MAXVAR], f[strict::MAXVAR];
    for (int l = 0; l < 6; l++) {
      for (int v = 0; v < 5; v++) s[l][v] = load_fluid<GEN>(P, pt[l], v);
      for (int v = 0; v < P.nchem; v++) {
        const double x = (GEN && pt[l].src >= 0) ? P.ghost[pt[l].src].buf[pt[l].off + 5 + v]
                                                 : P.w[5][(long)pt[l].off * P.nchem + v];
        s[l][5 + v] = (GEN && ((pt[l].neg >> 5) & 1u)) ? -x : x;
      }
    }
    const double p3 = strict::eos(P.gamma, s[3][0], s[3][1], s[3][2], s[3][3], s[3][4]);
    const int sbits = ((s[3][0] > 0.0) ? 0 : 1) | ((s[3][4] > 0.0) ? 0 : 2) | ((p3 > 0.0) ? 0 : 4);
    strict::face_flux(s, nvar, dir, P.gamma, f);
    if (PART != PART_TRACERS)
      for (int v = 0; v < 5; v++) emit(v, f[v]);
    if (PART != PART_FLUID && P.nchem > 0) {
      emit.species_begin();
      for (int v = 0; v < P.nchem; v += 2) emit.pair_next(f[5 + v], (v + 1 < P.nchem) ? f[6 + v] : 0.0, v + 1 < P.nchem);
    }
    (void)fn; (void)f1; (void)f2;
    return (PART != PART_TRACERS) ? sbits : 0;
  }
#else
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
    if (EB_ABLATE & 16) {
#pragma unroll
      for (int l = 0; l < 6; l++) { s.rinv[l] = 0.6 + 1e-9 * (double)pt[l].off; s.p[l] = 0.9 + 1e-9 * (double)pt[l].off; s.c[l] = 1.0 + 1e-9 * (double)pt[l].off; }
      s.srL = 1.2 + 1e-9 * (double)pt[2].off;
      s.srR = 1.2 + 1e-9 * (double)pt[3].off;
    } else if (!GEN && P.aux[0] != nullptr) {
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
    // sign of the species values of each point (Dirichlet ghosts negate them)
    double sg[6];
    bool owned = true;         // no point of this stencil lies in a halo slab
#pragma unroll
    for (int l = 0; l < 6; l++) {
      sg[l] = (GEN && ((pt[l].neg >> 5) & 1u)) ? -1.0 : 1.0;
      if (GEN) owned = owned && (pt[l].src < 0);
    }
    // Species two at a time from the pair-interleaved copy: one coalesced 16-byte load per stencil
    // point feeds two independent reconstructions (four WENO chains in flight), and the next pair
    // is loaded while the current one is reconstructed.  Stencils that reach into a halo slab
    // (reference wire layout, NVAR values per cell) take the scalar path below, as does a launch
    // without the copy.
    // 16-byte loads of species pairs: from the pair-interleaved copy (chemT), or from the vector
    // itself when nchem is even and its base 16-byte aligned (pair q of cell c is then the double2
    // number c nchem/2 + q)
    const bool vec_aos = P.chemT == nullptr && (P.nchem & 1) == 0 && (((unsigned long long)P.w[5]) & 15ull) == 0ull;
    if ((P.chemT != nullptr || vec_aos) && owned) {
      const long N = P.nx * P.ny * P.nz;
      const int npf = P.nchem / 2;              // full pairs; an odd species count leaves one behind
      // pair q of stencil point l is at qb[io[l]]: one 64-bit base for all six points, advanced once
      // per pair, and 32-bit offsets (one IMAD.WIDE per load)
      const double2* qb = STG ? sb : reinterpret_cast<const double2*>(vec_aos ? P.w[5] : P.chemT);
      const unsigned istride = vec_aos ? (unsigned)npf : 1u;
      const long qstep = (STG || vec_aos) ? 1 : N;
      unsigned io[6];
#pragma unroll
      for (int l = 0; l < 6; l++) io[l] = STG ? (unsigned)(l * ss) : pt[l].off * istride;
      // Slots 0 .. npf-1 hold two species, slot npf (odd nchem, chemT only) one.  The next slot is
      // loaded while the current one is reconstructed; the loop body is branch-free on purpose (the
      // last iteration re-loads its own slot: a conditional load here makes ptxas spill ~4 KB).
      const int ns = npf + (P.nchem & 1);
      double2 c[6], cn[6];
#pragma unroll
      for (int l = 0; l < 6; l++) {
        if (EB_ABLATE & 2) { c[l].x = up[l] * 1.25 + (double)pt[l].off * 1e-9; c[l].y = um[l] * 0.75; }
        else c[l] = qb[io[l]];
      }
      emit.species_begin();
#pragma unroll 1
      for (int q = 0; q < ns; q++) {
        qb += (q + 1 < ns) ? qstep : 0;
#pragma unroll
        for (int l = 0; l < 6; l++) {
          if (EB_ABLATE & 2) { cn[l].x = c[l].y * 1.0625 + (double)q; cn[l].y = c[l].x * 0.9375; }
          else cn[l] = qb[io[l]];
        }
        double a[6], b[6];
#pragma unroll
        for (int l = 0; l < 6; l++) {
          a[l] = GEN ? c[l].x * sg[l] : c[l].x;
          b[l] = GEN ? c[l].y * sg[l] : c[l].y;
        }
        const double fa = tracer_face(a, up, um);
        const double fb = tracer_face(b, up, um);
        emit.pair_next(fa, fb, q < npf);
#pragma unroll
        for (int l = 0; l < 6; l++) c[l] = cn[l];
      }
    } else {
      const double* cp[6];
#pragma unroll
      for (int l = 0; l < 6; l++) {
        if (GEN && pt[l].src >= 0) cp[l] = P.ghost[pt[l].src].buf + pt[l].off + 5;
        else cp[l] = P.w[5] + (long)pt[l].off * P.nchem;
      }
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
#endif  // EB_STRICT
}

template <bool AG, int PART, class Emit>
EB_HD int face_dispatch(bool gen, const RhsParams& P, int dir, int i, int j, int k, Emit emit)
{
  return gen ? face_all<true, AG, PART>(P, dir, i, j, k, emit) : face_all<false, AG, PART>(P, dir, i, j, k, emit);
}

// What face_all hands its fluxes to.  Fluid fields: emit(v, flux).  Species, in order and two at a
// time: emit.species_begin(), then emit.pair_next(fa, fb, two) per pair (two == false: the odd one
// out) -- the emitter keeps running pointers, so that the species loop carries no index
// arithmetic (on B200 an integer instruction next to the DFMA stream is not free, profiles/README.md).
// Array strides are compile-time constants in the instantiations with a fixed tile shape (S > 0).
// (1) into the shared-memory slot of the field (FX / FY / ZLO of rhs_fused_kernel)
template <int S>
struct EmitSlot {
  double* base;
  int rs;              // run-time stride (S == 0)
  double* sp;
  EB_HD int stride() const { return S > 0 ? S : rs; }
  EB_HD void operator()(int v, double x) const { base[v * stride()] = x; }
  EB_HD void species_begin() { sp = base + 5 * stride(); }
  EB_HD void pair_next(double a, double b, bool two) { sp[0] = a; if (two) sp[stride()] = b; sp += 2 * stride(); }
};
// (2) the flux through the z-face above the cell closes the divergence of its field (sum order of
// utilities.cpp:202-207) and the result goes out to wdot; a species pair as one 16-byte store
// (TRXc: stride of FX where it differs from ZLO's; XU: the flux through the upper x-face is read through the
// separate pointer FXU instead of from the next slot, rhs_fused_kernel<..., XC>)
template <bool GW, int TRc, int Tc, int TRXc = TRc, bool XU = false>
struct EmitDiv {
  const RhsParams& P;
  double* FX;
  const double* FXU;
  const double* fxu;
  double* FY;
  double* ZLO;
  int rTR, rT, TX;
  long cell;
  double *fx, *fy, *zl, *dst;    // running pointers of the species part
  bool fast;                     // species pairs: plain 16-byte stores
  EB_HD int TR() const { return TRc > 0 ? TRc : rTR; }
  EB_HD int TRX() const { return TRXc > 0 ? TRXc : rTR; }
  EB_HD int T() const { return Tc > 0 ? Tc : rT; }
  EB_HD double close(int v, double zup) const
  {
#if defined(EB_STRICT) || defined(EB_TRUE_DIVISION)
    const double div = (((XU ? FXU[v * TRX()] : FX[v * TRX() + 1]) - FX[v * TRX()]) / P.dx + (FY[v * T() + TX] - FY[v * T()]) / P.dy)
                       + (zup - ZLO[v * TR()]) / P.dz;
#else
    const double div = (((XU ? FXU[v * TRX()] : FX[v * TRX() + 1]) - FX[v * TRX()]) * P.rdx + (FY[v * T() + TX] - FY[v * T()]) * P.rdy)
                       + (zup - ZLO[v * TR()]) * P.rdz;
#endif
    ZLO[v * TR()] = zup;
    return div;
  }
  EB_HD void store(int v, double div) const
  {
    double* d = (v < 5) ? P.wdot[v] + cell : P.wdot[5] + cell * P.nchem + (v - 5);
    // GW: the caller ran external_forces itself, wdot already holds G (utilities.cpp:65)
    st_out(d, (GW ? *d : (v < 5 ? P.forcing[v] : 0.0)) - div);
  }
  EB_HD void operator()(int v, double zup) const { store(v, close(v, zup)); }
  EB_HD void species_begin()
  {
    fx = FX + 5 * TRX(); fy = FY + 5 * T(); zl = ZLO + 5 * TR();
    if (XU) fxu = FXU + 5 * TRX();
    dst = P.wdot[5] + cell * P.nchem;
    fast = !GW && P.vec_store;
  }
  EB_HD double close_next(int q, double zup) const      // q-th species (0 / 1) at the running pointers
  {
#if defined(EB_STRICT) || defined(EB_TRUE_DIVISION)
    return (((XU ? fxu[q * TRX()] : fx[q * TRX() + 1]) - fx[q * TRX()]) / P.dx + (fy[q * T() + TX] - fy[q * T()]) / P.dy)
           + (zup - zl[q * TR()]) / P.dz;
#else
    return (((XU ? fxu[q * TRX()] : fx[q * TRX() + 1]) - fx[q * TRX()]) * P.rdx + (fy[q * T() + TX] - fy[q * T()]) * P.rdy)
           + (zup - zl[q * TR()]) * P.rdz;
#endif
  }
#if !defined(EB_STRICT) && !defined(EB_TRUE_DIVISION)
  // minus the divergence, formed with the differences the other way round: bit for bit 0 - close_next() but
  // for the sign of a zero, without the subtraction from zero
  EB_HD double close_next_neg(int q, double zup) const
  {
    return ((fx[q * TRX()] - (XU ? fxu[q * TRX()] : fx[q * TRX() + 1])) * P.rdx + (fy[q * T()] - fy[q * T() + TX]) * P.rdy)
           + (zl[q * TR()] - zup) * P.rdz;
  }
#endif
  EB_HD void pair_next(double za, double zb, bool two)    // two == false: the odd species out, zb unused
  {
#if !defined(EB_STRICT) && !defined(EB_TRUE_DIVISION)
    if (fast) {                         // (nchem even: always two)
      const double na = close_next_neg(0, za), nb = close_next_neg(1, zb);
      zl[0] = za;
      zl[TR()] = zb;
      st_out2(dst, na, nb);
      fx += 2 * TRX(); fy += 2 * T(); zl += 2 * TR(); dst += 2;
      if (XU) fxu += 2 * TRX();
      return;
    }
#endif
    const double da = close_next(0, za);
    zl[0] = za;
    if (fast) {                         // (nchem even: always two)
      const double db = close_next(1, zb);
      zl[TR()] = zb;
      st_out2(dst, 0.0 - da, 0.0 - db);
    } else {
      st_out(dst, (GW ? dst[0] : 0.0) - da);
      if (two) {
        const double db = close_next(1, zb);
        zl[TR()] = zb;
        st_out(dst + 1, (GW ? dst[1] : 0.0) - db);
      }
    }
    fx += 2 * TRX(); fy += 2 * T(); zl += 2 * TR(); dst += 2;
    if (XU) fxu += 2 * TRX();
  }
};

#if defined(__CUDACC__) || defined(EB_CUDA_EMU)

// Pre-pass over the cells [c0, c1): per-cell 1/rho, p, c, sqrt(rho) (40 B read, 32 B written per
// cell), so that the 18 stencils a cell sits in do not each redo a reciprocal and two square roots on
// the FP64 pipe; and the pair-interleaved copy of the species (8 nchem B read and written per cell;
// see RhsParams::chemT), read with one thread per 16-byte chunk so that both sides stay coalesced.
__global__ void aux_kernel(const RhsParams P, double* a0, double* a1, double* a2, double* a3, double* chemT,
                           long c0, long c1)
{
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
  for (long c = c0 + tid; c < c1; c += nthr) {
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
  if (chemT != nullptr && P.nchem > 0) {
    // blocks of 256 cells per CTA; chunk j of a block is species pair j % np of its cell j / np
    // (32-bit arithmetic: a 64-bit division per chunk would cost as much as the copy itself)
    const long N = P.nx * P.ny * P.nz;
    const unsigned np = (unsigned)(P.nchem + 1) / 2u;
    const bool vec = (P.nchem & 1) == 0 && (((unsigned long long)P.w[5]) & 15ull) == 0ull;
    const double2* src2 = reinterpret_cast<const double2*>(P.w[5]);
    double2* dst2 = reinterpret_cast<double2*>(chemT);
    const long nblk = (c1 - c0 + 255) / 256;
    for (long b = blockIdx.x; b < nblk; b += gridDim.x) {
      const long cb = c0 + b * 256;
      const unsigned ncb = (unsigned)((c1 - cb < 256) ? c1 - cb : 256);
      for (unsigned j = threadIdx.x; j < ncb * np; j += blockDim.x) {
        const unsigned cl = j / np, h = j - cl * np;
        const long c = cb + cl;
        if (vec) {
          dst2[(long)h * N + c] = src2[cb * np + j];
        } else {
          double2 x;
          x.x = P.w[5][c * P.nchem + 2 * h];
          x.y = (2 * (int)h + 1 < P.nchem) ? P.w[5][c * P.nchem + 2 * h + 1] : 0.0;
          dst2[(long)h * N + c] = x;
        }
      }
    }
  }
}

// fslow / fexpl post-step (multirate_chem_hydro_main.cpp:1059-1068, imex_chem_hydro_main.cpp fexpl):
// chemdot[nchem-1] = etdot, etdot = 0 for the cells of the box [P.lo, P.hi) -- 8 B read, 16 B written
// per cell, on the device (the reference copies the whole chemistry vector to the host and back
// around fEuler in its GPU builds, :1028-1031,1070-1073).
__global__ void slow_post_kernel(const RhsParams P)
{
  const long ex = P.hi[0] - P.lo[0], ey = P.hi[1] - P.lo[1], ez = P.hi[2] - P.lo[2];
  const long n = ex * ey * ez;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    const long k = q / (ex * ey), r = q - k * ex * ey, j = r / ex, i = r - j * ex;
    const long c = (P.lo[0] + i) + P.nx * ((P.lo[1] + j) + P.ny * (P.lo[2] + k));
    P.wdot[5][c * P.nchem + (P.nchem - 1)] = P.wdot[4][c];
    P.wdot[4][c] = 0.0;
  }
}

#if defined(__CUDACC__)
// named barrier: `count` threads (whole warps) meet at hardware barrier `id` (1..15).  The count is
// always 64 here (two neighbouring warp rows) and is given as an immediate: with a register operand
// compute-sanitizer's synccheck cannot see that the barrier is a partial one and reports the other
// rows as divergent.
__device__ __forceinline__ void eb_bar_sync(int id, int count)
{
  (void)count;
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// ---- bulk-copy staging (rhs_fused_kernel<..., STAGE>): one elected thread issues cp.async.bulk copies
// global -> shared (SASS: UBLKCP) that complete on an mbarrier; every thread waits on its phase.
__device__ __forceinline__ unsigned eb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void eb_mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(eb_smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void eb_mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(eb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void eb_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(eb_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(eb_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void eb_mbar_arrive(unsigned long long* bar)      // release.cta
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(eb_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void eb_mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "EB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra EB_DONE_%=;\n"
      "bra EB_WAIT_%=;\n"
      "EB_DONE_%=:\n"
      "}\n" ::"r"(eb_smem_addr(bar)), "r"(parity) : "memory");
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
// TYC: 0, or the number of tile rows of a launch whose CTAs are 32 x TYC threads (tile shape and
// shared-memory strides are then compile-time constants).
// STAGE (A/B variant, EULERB200_STAGE=1; needs TYC > 0, an even species count and CTA-wide barriers): tiles
// whose x- and y-stencils stay inside the box keep the species of the CURRENT plane -- tile plus
// three halo cells, (TX+5) x (TY+5) cells x nchem values, rows of the species-fastest vector are
// contiguous -- in shared memory, filled by cp.async.bulk copies that one thread issues for plane k+1
// while the z-faces of plane k are computed; the x- and y-face species loops then read shared memory
// (16-byte loads at stride 8 nchem B: conflict-free per quarter warp for nchem = 10) instead of 19
// cache lines per load.  The z-stencil (six planes) cannot be staged: 6 x tile x 8 nchem B = 184 KB.
template <int MAXT, int MINB, bool GW = false, bool AG = false, int PART = PART_ALL, int TYC = 0, bool STAGE = false, bool XC = false>
__global__ void __launch_bounds__(MAXT, MINB) rhs_fused_kernel(const RhsParams P)
{
  static_assert(!XC || (TYC > 1 && !STAGE), "XC needs a compiled-in tile of 32 x TYC threads");
  EB_DYN_SMEM(double, smem);
  const int TX = TYC > 0 ? 32 : (int)blockDim.x, TY = TYC > 0 ? TYC : (int)blockDim.y, T = TX * TY;
  constexpr int Tc = 32 * TYC, TRc = TYC > 0 ? 32 * (TYC - 1) : 0;
  constexpr int TRXc = XC ? 34 * (TYC - 1) : TRc;        // XC: FX rows carry two more slots, the face column of the top warp (two buffers)
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * TX + tx;
  // fields of this launch: v0 .. v0+nf-1 in the reference's order
  const int v0 = (PART == PART_TRACERS) ? 5 : 0;
  const int nf = (PART == PART_ALL) ? 5 + P.nchem : (PART == PART_FLUID ? 5 : P.nchem);
  const bool pair = !STAGE && P.pair_sync != 0;          // (STAGE re-fills the staged plane behind a CTA-wide barrier)
  const bool two_fy = !STAGE && P.pair_sync == 1;
  // the face-only top row of the tile never touches FX / ZLO: those arrays are [NF][TR]
  const int TR = T - TX;
  const int TRX = XC ? TRXc : TR;

  const int nx = (int)P.nx, ny = (int)P.ny, nz = (int)P.nz;
  const int hix = (int)P.hi[0], hiy = (int)P.hi[1], hiz = (int)P.hi[2];
  const int ti0 = (int)P.lo[0] + (int)blockIdx.x * (XC ? TX : TX - 1);
  const int tj0 = (int)P.lo[1] + (int)blockIdx.y * (TY - 1);
  const int i = ti0 + tx, j = tj0 + ty;
  const int k0 = (int)P.lo[2] + (int)blockIdx.z * P.seg_len;
  const int k1 = (k0 + P.seg_len < hiz) ? k0 + P.seg_len : hiz;

  const bool row_ok = (ty < TY - 1) && (j < hiy);
  const bool col_ok = (XC || tx < TX - 1) && (i < hix);
  const bool owns = row_ok && col_ok;                 // this thread owns a cell column
  bool need_x = row_ok && (i <= hix);                 // lower x-face at position i
  const bool need_y = col_ok && (j <= hiy);           // lower y-face at position j

  // XC: the tile owns all 32 columns; the x-faces that close it on the right (position ti0 + 32,
  // rows 0 .. TY-2) are computed by lanes 0 .. TY-2 of the top warp -- whose own row only supplies
  // y-faces -- into slot 32 + b of each FX row, b = plane parity (a tile cut short by the box closes
  // itself: the lane at i == hix computes that face as before).  The top warp works one plane ahead
  // of the rows (the loop starts one step early for it) and hands each column over through a pair of
  // mbarriers per buffer (xfull[b]: column written; xfree[b]: every row has consumed it), so that
  // neither side normally waits and the rows stay as loosely coupled as the row rendezvous leaves them.
  const bool top = XC && ty == TY - 1;
  const bool xc_tile = XC && (ti0 + TX <= hix);        // CTA-uniform
  int xi = i, xj = j, xslot = XC ? ty * 34 + tx : t;
  if (top) {
    xi = ti0 + TX; xj = tj0 + tx;
    need_x = xc_tile && tx < TY - 1 && xj < hiy;
    xslot = (tx < TY - 1 ? tx : 0) * 34 + 32;
  }
  double* FX = smem + xslot - (long)v0 * TRX;           // indexed with the global field number v
  double* FY = smem + (long)nf * TRX + t - (long)v0 * T;
  long fy_flip = two_fy ? (long)nf * T : 0;         // signed distance to the other FY buffer
  double* ZLO = smem + (long)nf * (TRX + (two_fy ? 2L : 1L) * T) + t - (long)v0 * TR;
  unsigned long long* xfull = reinterpret_cast<unsigned long long*>(smem + (long)nf * (TRX + (two_fy ? 2L : 1L) * T + TR));
  unsigned long long* xfree = xfull + 2;
  if (XC) {
    if (t == 0) { eb_mbar_init(xfull, 1); eb_mbar_init(xfull + 1, 1); eb_mbar_init(xfree, TY - 1); eb_mbar_init(xfree + 1, TY - 1); }
    __syncthreads();
  }

  // CTA-uniform: does any stencil of this tile reach beyond the owned range in x / y?
  const bool gen_x = (ti0 - 3 < 0) || (ti0 + (XC ? TX : TX - 1) + 2 >= nx);
  const bool gen_y = (tj0 - 3 < 0) || (tj0 + TY - 1 + 2 >= ny);

  int mask = 0;

  // ---- STAGE: shared-memory copy of the species of the current plane
  const int SW = TX + 5;                                   // cells per staged row
  double2* stage = nullptr;
  unsigned long long* mbar = nullptr;
  bool staged = false;
  unsigned sphase = 0;
  if (STAGE) {
    double* after = smem + (long)nf * (2L * TR + T);        // (single FY buffer: STAGE runs with CTA-wide barriers)
    stage = reinterpret_cast<double2*>(after);
    mbar = reinterpret_cast<unsigned long long*>(after + (long)SW * (TY + 5) * P.nchem);
    staged = !gen_x && !gen_y && P.nchem > 0;              // CTA-uniform
    if (staged) {
      if (t == 0) eb_mbar_init(mbar, 1);
      __syncthreads();
    }
  }
  auto stage_plane = [&](int k) {                          // thread 0: species rows tj0-3 .. tj0+TY+1 of plane k
    const unsigned row_bytes = (unsigned)(SW * P.nchem * sizeof(double));
    eb_mbar_expect_tx(mbar, row_bytes * (unsigned)(TY + 5));
    for (int r = 0; r < TY + 5; r++)
      eb_bulk_g2s(reinterpret_cast<double*>(stage) + (long)r * SW * P.nchem,
                  P.w[5] + ((long)(ti0 - 3) + (long)nx * ((tj0 - 3 + r) + (long)ny * k)) * P.nchem, row_bytes, mbar);
  };
  if (STAGE && staged && t == 0) stage_plane(k0);
  const int npf = P.nchem / 2;

  // z-face below the first plane of the segment
  if (owns)
    face_dispatch<AG, PART>(k0 - 3 < 0 || k0 + 2 >= nz, P, 2, i, j, k0, EmitSlot<TRc>{ZLO, TR, nullptr});

  for (int k = XC ? k0 - 1 : k0; k < k1; k++) {
    // (XC: in the leading step k0 - 1 only the top warp works -- the face column of plane k0)
    const bool act = !XC || k >= k0;
    const int xn = k - k0;                                 // XC: plane number within the segment
    // ---- phase A: lower x- and y-faces of plane k -> shared memory ----
    if (STAGE && staged) {
      eb_mbar_wait(mbar, sphase);                          // plane k has landed
      sphase ^= 1u;
      // species of cell (row r, column c) of the staged window: stage[(r SW + c) npf + q]
      if (need_x) {
        const int bits = face_all<false, AG, PART, EmitSlot<TRc>, true>(P, 0, i, j, k, EmitSlot<TRc>{FX, TR, nullptr},
                                                                       stage + ((ty + 3) * SW + tx) * npf, npf);
        if (owns) mask |= bits;
      }
      if (need_y)
        face_all<false, AG, PART, EmitSlot<Tc>, true>(P, 1, i, j, k, EmitSlot<Tc>{FY, T, nullptr},
                                                      stage + (ty * SW + tx + 3) * npf, SW * npf);
    } else if (XC) {
      // y-faces first: they are all the row rendezvous publishes (FX stays inside the warp, or comes
      // from the top warp through xfull)
      if (act && need_y)
        face_dispatch<AG, PART>(gen_y, P, 1, i, j, k, EmitSlot<Tc>{FY, T, nullptr});
      if (pair) {
        if (ty > 0) eb_bar_sync(ty, 64);
        if (ty < TY - 1) eb_bar_sync(ty + 1, 64);
      } else {
        __syncthreads();
      }
      // x-faces, one call site for both kinds of warp.  Rows: plane k.  Top warp: the column of plane k + 1
      // (number xn + 1) into buffer (xn + 1) & 1, once the rows are done with the column of plane k - 1 there.
      const int n1 = xn + 1;
      const bool tcol = top && xc_tile && k + 1 < k1;
      if (tcol && n1 >= 2) eb_mbar_wait(xfree + (n1 & 1), (unsigned)((n1 >> 1) - 1) & 1u);
      if (top ? (tcol && need_x) : (act && need_x)) {
        const int bits = face_dispatch<AG, PART>(gen_x, P, 0, xi, xj, top ? k + 1 : k,
                                                 EmitSlot<TRXc>{top ? FX + (n1 & 1) : FX, TRX, nullptr});
        if (owns) mask |= bits;
      }
      __syncwarp();                                        // FX of the warp's own lanes
      if (tcol) {
        if (tx == 0) eb_mbar_arrive(xfull + (n1 & 1));
      } else if (!top && act && xc_tile) {
        eb_mbar_wait(xfull + (xn & 1), (unsigned)(xn >> 1) & 1u);
      }
    } else {
    if (need_x) {
      const int bits = face_dispatch<AG, PART>(gen_x, P, 0, i, j, k, EmitSlot<TRc>{FX, TR, nullptr});
      if (owns) mask |= bits;
    }
    if (need_y)
      face_dispatch<AG, PART>(gen_y, P, 1, i, j, k, EmitSlot<Tc>{FY, T, nullptr});
    }
    if (XC) {
    } else if (EB_ABLATE & 1) {
    } else if (pair) {
      if (ty > 0) eb_bar_sync(ty, 64);
      if (ty < TY - 1) eb_bar_sync(ty + 1, 64);
    } else {
      __syncthreads();
    }

    // every thread is past its reads of the staged plane (CTA-wide barrier above): fetch the next one
    // behind the z-faces
    if (STAGE && staged && t == 0 && k + 1 < k1) stage_plane(k + 1);

    // ---- phase B: z-face above cell (i,j,k); each flux closes the divergence of its
    //      field as soon as it exists (sum order of utilities.cpp:202-207) ----
    if (owns && act) {
      const long cell = i + nx * (j + ny * k);
      // (XC: lane 31 finds the flux through its upper x-face in the column slot of this plane's parity)
      const double* FXU = FX + ((XC && tx == TX - 1 && xc_tile) ? 1 + (xn & 1) : 1);
      face_dispatch<AG, PART>(k + 1 - 3 < 0 || k + 1 + 2 >= nz, P, 2, i, j, k + 1,
                              EmitDiv<GW, TRc, Tc, TRXc, XC>{P, FX, FXU, nullptr, FY, ZLO, TR, T, TX, cell, nullptr, nullptr, nullptr, nullptr, false});
    }
    if (XC && xc_tile && !top && act) {     // this row is done with the column of plane k
      __syncwarp();
      if (tx == 0) eb_mbar_arrive(xfree + (xn & 1));
    }
    if (EB_ABLATE & 1) {
    } else if (two_fy) {
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
