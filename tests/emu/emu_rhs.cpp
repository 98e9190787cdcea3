// ---------------------------------------------------------------------------
// emu_rhs.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Runs the kernel source of the product (rhs_kernel.cuh) and its host-side set-up
// (host_setup.h) on the CPU through tests/emu/cuda_emu.h, for the "not gpu" tests.
// ---------------------------------------------------------------------------
#include "cuda_emu.h"
#include "../../sundials-manyvector-demo_b200/csrc/host_setup.h"
#include "../../sundials-manyvector-demo_b200/csrc/halo_kernels.cuh"
#include "../../sundials-manyvector-demo_b200/csrc/vector_kernels.cuh"

extern "C" int emu_rhs(const eulerb200_config* cfg, const double* const* w, double* const* wdot,
                       const double* const* recv, int* state_bits, const long* lo, const long* hi,
                       int threads, int use_aux, double energy_units, int pair, int g_in_wdot, int aux_in_gen, int split, int use_chemT, int stage, int xc)
{
  eb::RhsParams P;
  std::vector<double> aux[4];
  P.nx = cfg->nxl; P.ny = cfg->nyl; P.nz = cfg->nzl;
  P.nchem = cfg->nchem;
  P.gamma = cfg->gamma;
  P.rdx = EB_RD_SCALE / cfg->dx; P.rdy = EB_RD_SCALE / cfg->dy; P.rdz = EB_RD_SCALE / cfg->dz;
  P.dx = cfg->dx; P.dy = cfg->dy; P.dz = cfg->dz;
  for (int f = 0; f < 5; f++) P.forcing[f] = cfg->forcing[f];
  for (int f = 0; f < 6; f++) { P.w[f] = w[f]; P.wdot[f] = wdot[f]; }
  for (int f = 0; f < 6; f++)
    if (eb::ghost_face(*cfg, f, recv ? recv[f] : nullptr, &P.ghost[f]) != 0) return -1;
  for (int q = 0; q < 4; q++) P.aux[q] = nullptr;
  P.chemT = nullptr;
  std::vector<double> chemT;
  P.slow_mode = 0; P.inv_energy_units = 1.0; P.et_rw = nullptr;
  P.vec_store = (cfg->nchem > 0 && (cfg->nchem & 1) == 0 && (((unsigned long long)wdot[5]) & 15ull) == 0ull) ? 1 : 0;
  if (energy_units > 0.0) {        // fslow mode: aux_kernel rebuilds the total energy in w[4]
    P.slow_mode = 1;
    P.inv_energy_units = 1.0 / energy_units;
    P.et_rw = const_cast<double*>(w[4]);
  }
  if (use_chemT && P.nchem > 0) {
    chemT.assign((size_t)2 * ((P.nchem + 1) / 2) * P.nx * P.ny * P.nz, 0.0 / 0.0);
    P.chemT = chemT.data();
  }
  if (use_aux || P.slow_mode || P.chemT) {    // the product's own pre-pass kernel, launched like launch_aux() launches it
    const long N = P.nx * P.ny * P.nz;
    if (use_aux) for (int q = 0; q < 4; q++) aux[q].assign(N, 0.0 / 0.0);
    double* a[4] = {nullptr, nullptr, nullptr, nullptr};
    if (use_aux) for (int q = 0; q < 4; q++) a[q] = aux[q].data();
    cuda_emu::launch_plain(eb::aux_kernel, dim3((unsigned)std::min<long>((N + 255) / 256, 148L * 16)), dim3(256), P,
                           a[0], a[1], a[2], a[3], const_cast<double*>(P.chemT), 0L, N);
    if (use_aux) for (int q = 0; q < 4; q++) P.aux[q] = aux[q].data();
  }
  int flag = 0;
  P.state_flag = &flag;
  const long full_lo[3] = {0, 0, 0}, full_hi[3] = {P.nx, P.ny, P.nz};
  for (int d = 0; d < 3; d++) { P.lo[d] = lo ? lo[d] : full_lo[d]; P.hi[d] = hi ? hi[d] : full_hi[d]; }
  // the launches launch_box() makes on the device: one fused launch, or (split) fluid then species
  const int parts[2] = {split && P.nchem > 0 ? eb::PART_FLUID : eb::PART_ALL, eb::PART_TRACERS};
  for (int q = 0; q < (split && P.nchem > 0 ? 2 : 1); q++) {
    const int part = parts[q];
    const int nf = part == eb::PART_ALL ? 5 + P.nchem : (part == eb::PART_FLUID ? 5 : P.nchem);
    eb::LaunchGeom L = eb::launch_geom(P.lo, P.hi, nf, threads, pair);
    // the XC instantiations exist for the compiled-in tile shapes only (launch_part())
    if (xc && L.tx == 32 && (L.ty == 12 || L.ty == 4)) L = eb::launch_geom(P.lo, P.hi, nf, threads, pair, 5920, 1);
    P.pair_sync = L.pair;
    if (pair != 0 && !L.pair) return -77;           // rows are not warps: the pairwise path does not apply
    P.seg_len = L.seg_len;
    const dim3 grid(L.gx, L.gy, L.gz), block(L.tx, L.ty, 1);
#define EMU_PARTS_X(GW, AG, TYC, XC)                                                                                   \
    do {                                                                                                              \
      if (part == eb::PART_ALL) cuda_emu::launch(eb::rhs_fused_kernel<256, 1, GW, AG, eb::PART_ALL, TYC, false, XC>, grid, block, L.smem, P);      \
      else if (part == eb::PART_FLUID) cuda_emu::launch(eb::rhs_fused_kernel<256, 1, GW, AG, eb::PART_FLUID, TYC, false, XC>, grid, block, L.smem, P); \
      else cuda_emu::launch(eb::rhs_fused_kernel<256, 1, GW, AG, eb::PART_TRACERS, TYC, false, XC>, grid, block, L.smem, P);                \
    } while (0)
#define EMU_PARTS(GW, AG, TYC) EMU_PARTS_X(GW, AG, TYC, false)
    // like launch_part(): the instantiation with the tile shape compiled in where there is one (here
    // 32 x 4 and 32 x 12), else the any-shape one
#define EMU_LAUNCH(GW, AG)                                                                                            \
    do {                                                                                                              \
      if (L.xc && L.ty == 12) EMU_PARTS_X(GW, AG, 12, true);                                                          \
      else if (L.xc && L.ty == 4) EMU_PARTS_X(GW, AG, 4, true);                                                       \
      else if (L.tx == 32 && L.ty == 12) EMU_PARTS(GW, AG, 12);                                                       \
      else if (L.tx == 32 && L.ty == 4) EMU_PARTS(GW, AG, 4);                                                         \
      else EMU_PARTS(GW, AG, 0);                                                                                      \
    } while (0)
    // the bulk-copy staging variant (launch_part(): default kind, fused launch, 32 x 12 tile, CTA-wide barriers)
    if (stage && !g_in_wdot && !aux_in_gen && part == eb::PART_ALL && L.tx == 32 && L.ty == 12 && P.vec_store && !P.chemT) {
      L = eb::launch_geom(P.lo, P.hi, nf, threads, 0);
      P.pair_sync = 0;
      L.smem += sizeof(double) * (size_t)(L.tx + 5) * (L.ty + 5) * P.nchem + 16;
      cuda_emu::launch(eb::rhs_fused_kernel<256, 1, false, false, eb::PART_ALL, 12, true>, dim3(L.gx, L.gy, L.gz), dim3(L.tx, L.ty, 1), L.smem, P);
      continue;
    }
    // the three instantiations launch_box() chooses from on the device
    if (g_in_wdot) EMU_LAUNCH(true, false);
    else if (aux_in_gen) EMU_LAUNCH(false, true);
    else EMU_LAUNCH(false, false);
#undef EMU_LAUNCH
#undef EMU_PARTS
#undef EMU_PARTS_X
  }
  if (P.slow_mode) cuda_emu::launch_plain(eb::slow_post_kernel, dim3(3), dim3(64), P);   // as launch_box() does
  *state_bits = flag;
  return flag ? -1 : 0;
}

extern "C" int emu_decompose(int nprocs, int rank, const int64_t* n, const int32_t* bc,
                             int32_t* dims, int32_t* coords, int64_t* ext, int32_t* nbr)
{
  return eb::decompose(nprocs, rank, n, bc, dims, coords, ext, nbr);
}

// The interior / shell boxes rhs_impl() launches for a rank with the given remote faces (host_setup.h), with the
// tile pitches of the launch geometry of the whole box.  out: count, then lo[3], hi[3] per box.
extern "C" int emu_overlap_boxes(const long* n, const int* remote, int nchem, int threads, int xc, int thick, long* out)
{
  const long z[3] = {0, 0, 0};
  const eb::LaunchGeom L = eb::launch_geom(z, n, 5 + nchem, threads, 2, 5920, xc);
  const long pitch[3] = {L.xc ? L.tx : L.tx - 1, L.ty - 1, L.seg_len};
  bool rem[6];
  for (int f = 0; f < 6; f++) rem[f] = remote[f] != 0;
  eb::BoxList B;
  const bool ok = eb::overlap_boxes(n, rem, pitch, thick != 0, &B);
  out[0] = ok ? B.count : 0;
  for (int q = 0; q < B.count; q++)
    for (int d = 0; d < 3; d++) { out[1 + 6 * q + d] = B.lo[q][d]; out[1 + 6 * q + 3 + d] = B.hi[q][d]; }
  return ok ? 0 : 1;
}

extern "C" double emu_boundary_tile_fraction(const long* lo, const long* hi, long nx, long ny, int nchem, int threads)
{
  const eb::LaunchGeom L = eb::launch_geom(lo, hi, 5 + nchem, threads, 2);
  return eb::boundary_tile_fraction(lo, hi, nx, ny, L);
}

// The face kernels of the halo exchange (halo_kernels.cuh), launched like exchange_start() /
// eulerb200_ghost_face() launch them.  what = 0: pack the send buffer of face f; 1: the ghost layers
// of face f (recv: this rank's halo slabs, NULL where a face has no remote neighbour).
extern "C" int emu_face(const eulerb200_config* cfg, const double* const* w, const double* const* recv, int f, int what,
                        double* out)
{
  eb::FaceGeom g;
  g.nx = cfg->nxl; g.ny = cfg->nyl; g.nz = cfg->nzl;
  g.nchem = cfg->nchem; g.f = f;
  for (int q = 0; q < 6; q++) g.w[q] = (q < 5 || cfg->nchem > 0) ? w[q] : nullptr;
  const long nent = eb::face_len(*cfg, f) / (5 + cfg->nchem);
  const dim3 grid((unsigned)((nent + 255) / 256)), block(256);
  if (what == 0) {
    cuda_emu::launch2(eb::pack_face_kernel, grid, block, g, out, nent);
  } else if (what == 2) {        // the peer-store transport's pack, as exchange_start() launches it
    cuda_emu::launch2(eb::pack_face_warp_kernel, dim3((unsigned)((nent + 7) / 8)), dim3(32, 8), g, out, nent);
  } else {
    eb::GhostFace G;
    if (eb::ghost_face(*cfg, f, recv ? recv[f] : nullptr, &G) != 0) return -1;
    cuda_emu::launch3(eb::ghost_face_kernel, grid, block, g, G, out, nent);
  }
  return 0;
}

// The reduction / vector kernels (vector_kernels.cuh), launched like the C ABI launches them.  They
// use barriers and warp shuffles, so they go through the fibre scheduler (one parameter struct each).
namespace {
struct WaveArgs { const double* w[5]; long N; double gamma; unsigned long long* out; };
void wave_entry(const WaveArgs a) { eb::wavespeed_kernel(a.w[0], a.w[1], a.w[2], a.w[3], a.w[4], a.N, a.gamma, a.out); }
struct LinArgs { eb::LinCombArgs a; double* out; long n; };
void lin_entry(const LinArgs a) { eb::lincomb_kernel(a.a, a.out, a.n); }
struct WrmsArgs { const double* x; const double* y; double rtol, atol; long n; double* acc; };
void wrms_entry(const WrmsArgs a) { eb::wrms_kernel(a.x, a.y, a.rtol, a.atol, a.n, a.acc); }
// (two warps per CTA and at most three CTAs: enough to exercise the warp, CTA and grid levels of the
// reductions while keeping the fibre count per launch small)
unsigned emu_blocks(long n) { return n <= 512 ? 1u : (n <= 4096 ? 2u : 3u); }
}  // namespace

extern "C" double emu_max_wavespeed(const double* const* w, long N, double gamma)
{
  unsigned long long bits = 0;
  WaveArgs a;
  for (int f = 0; f < 5; f++) a.w[f] = w[f];
  a.N = N; a.gamma = gamma; a.out = &bits;
  cuda_emu::launch(wave_entry, dim3(emu_blocks(N)), dim3(64), 0, a);
  double alpha;
  memcpy(&alpha, &bits, sizeof alpha);
  return alpha;
}
extern "C" void emu_lincomb(int nterms, const double* coef, const double* const* x, double* out, long n)
{
  LinArgs a;
  a.a.nterms = nterms;
  for (int t = 0; t < nterms; t++) { a.a.c[t] = coef[t]; a.a.x[t] = x[t]; }
  a.out = out; a.n = n;
  cuda_emu::launch_plain(lin_entry, dim3(emu_blocks(n)), dim3(64), a);      // no barriers or shuffles: no fibres needed
}
extern "C" void emu_wrms_accum(const double* x, const double* y, double rtol, double atol, long n, double* acc)
{
  WrmsArgs a = {x, y, rtol, atol, n, acc};
  cuda_emu::launch(wrms_entry, dim3(emu_blocks(n)), dim3(64), 0, a);
}

// One face through the product's face arithmetic (euler_math.cuh: cell_aux, fluid_face, tracer_face) on the
// reference's own face_flux interface: w1d[6][nvar] in the reference's field order, idir, f_face[nvar]
// (utilities.cpp:270).  The stencil is put into sweep-aligned order and the fluxes back exactly as face_all does.
extern "C" void emu_face_flux(const double* w1d, int nvar, int idir, double gamma, double* f_face)
{
#ifndef EB_STRICT
  const int fn = 1 + idir, f1 = (idir == 1) ? 1 : 2, f2 = (idir == 2) ? 1 : 3;
  eb::FluidStencil s;
  for (int l = 0; l < 6; l++) {
    const double* c = w1d + (long)l * nvar;
    s.r[l] = c[0]; s.mn[l] = c[fn]; s.m1[l] = c[f1]; s.m2[l] = c[f2]; s.e[l] = c[4];
    const eb::CellAux a = eb::cell_aux(gamma, s.r[l], s.mn[l], s.m1[l], s.m2[l], s.e[l]);
    s.rinv[l] = a.rinv; s.p[l] = a.p; s.c[l] = a.c;
    if (l == 2) s.srL = a.sr;
    if (l == 3) s.srR = a.sr;
  }
  double f[5], alpha, u[6];
  eb::fluid_face(s, gamma, f, alpha, u);
  const double half = EB_FOLD_HALF ? 0.5 : 1.0;       // the faces hand out twice the flux where the divergence halves it
  f_face[0] = half * f[0]; f_face[fn] = half * f[1]; f_face[f1] = half * f[2]; f_face[f2] = half * f[3]; f_face[4] = half * f[4];
  double up[6], um[6];
  for (int l = 0; l < 6; l++) { up[l] = u[l] + alpha; um[l] = u[l] - alpha; }
  for (int v = 5; v < nvar; v++) {
    double c[6];
    for (int l = 0; l < 6; l++) c[l] = w1d[(long)l * nvar + v];
    f_face[v] = half * eb::tracer_face(c, up, um);
  }
#else
  (void)w1d; (void)nvar; (void)idir; (void)gamma; (void)f_face;
#endif
}
