// ---------------------------------------------------------------------------
// emu_abi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE, never shipped or loaded by the product.
//
// The C-ABI entry points (include/eulerb200.h) that the drop-in fEuler / stability of
// sundials-manyvector-demo_b200/host/feuler_dropin.cpp and the native driver host/euler3d_b200.cpp
// call, implemented on top of the CPU emulation of the kernel SOURCE (emu_rhs.cpp), single rank.  It exists so that the "not gpu"
// tier can link the drop-in against the reference's own EulerData and run it next to the
// unmodified reference fEuler (oracle/dropin_check.cpp) without a GPU: that checks the host
// logic of the drop-in (probe of the external_forces hook, per-call hook, pointer plumbing,
// error paths), nothing about the device library.
// ---------------------------------------------------------------------------
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include "../../include/eulerb200.h"

extern "C" int emu_rhs(const eulerb200_config* cfg, const double* const* w, double* const* wdot,
                       const double* const* recv, int* state_bits, const long* lo, const long* hi,
                       int threads, int use_aux, double energy_units, int pair, int g_in_wdot, int aux_in_gen, int split, int use_chemT, int stage, int xc);

extern "C" double emu_max_wavespeed(const double* const* w, long N, double gamma);
extern "C" void emu_lincomb(int nterms, const double* coef, const double* const* x, double* out, long n);
extern "C" void emu_wrms_accum(const double* x, const double* y, double rtol, double atol, long n, double* acc);

struct eulerb200_ctx { eulerb200_config cfg; bool gw; std::string err; long launches; };
static std::string g_err;

extern "C" {
int eulerb200_version(void) { return EULERB200_VERSION; }
int eulerb200_create(const eulerb200_config* cfg, eulerb200_ctx** out)
{
  if (!cfg || !out || cfg->nranks != 1) { g_err = "emu_abi: single rank only"; return -1; }
  *out = new eulerb200_ctx{*cfg, false, "", 0};
  return 0;
}
int eulerb200_destroy(eulerb200_ctx* c) { delete c; return 0; }
const char* eulerb200_last_error(const eulerb200_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
int eulerb200_set_forcing_in_wdot(eulerb200_ctx* c, int32_t on) { if (!c) return -1; c->gw = on != 0; return 0; }
int eulerb200_rhs_any(eulerb200_ctx* c, double, const double* const* w, double* const* wdot, void*)
{
  int bits = 0;
  const int rc = emu_rhs(&c->cfg, w, wdot, nullptr, &bits, nullptr, nullptr, 128, 1, 0.0, 0, c->gw ? 1 : 0, 1, getenv("EULERB200_SPLIT") ? atoi(getenv("EULERB200_SPLIT")) : 0, 0, 0, getenv("EULERB200_XC") ? atoi(getenv("EULERB200_XC")) : 1);
  if (rc) c->err = "STATE_ERROR: legal_state (fEuler) failed with flag = " + std::to_string(bits);
  return rc;
}
int eulerb200_stability_any(eulerb200_ctx* c, const double* const* w, double cfl, double* dt_stab, void*)
{
  // utilities.cpp:505-520 through the product's wavespeed_kernel
  const long N = c->cfg.nxl * c->cfg.nyl * c->cfg.nzl;
  const double alpha = emu_max_wavespeed(w, N, c->cfg.gamma);
  *dt_stab = cfl * std::fmin(std::fmin(c->cfg.dx, c->cfg.dy), c->cfg.dz) / alpha;
  return 0;
}
// what the native driver (host/euler3d_b200.cpp) needs on top: "device" memory is host memory here
int eulerb200_rhs(eulerb200_ctx* c, double t, const double* const* w, double* const* wdot, void* s)
{
  c->launches += 2;
  return eulerb200_rhs_any(c, t, w, wdot, s);
}
int eulerb200_stability(eulerb200_ctx* c, const double* const* w, double cfl, double* dt_stab, void* s)
{
  return eulerb200_stability_any(c, w, cfl, dt_stab, s);
}
void* eulerb200_device_alloc(int64_t bytes) { return bytes > 0 ? malloc((size_t)bytes) : nullptr; }
void eulerb200_device_free(void* p) { free(p); }
int eulerb200_copy_to_device(void* dst, const void* src, int64_t bytes) { memcpy(dst, src, (size_t)bytes); return 0; }
int eulerb200_copy_to_host(void* dst, const void* src, int64_t bytes) { memcpy(dst, src, (size_t)bytes); return 0; }
int64_t eulerb200_launch_count(const eulerb200_ctx* c) { return c ? c->launches : -1; }
void* eulerb200_managed_alloc(int64_t bytes) { return bytes > 0 ? malloc((size_t)bytes) : nullptr; }
int eulerb200_synchronize(eulerb200_ctx* c) { return c ? 0 : -1; }
int eulerb200_profile(eulerb200_ctx* c, int32_t, int32_t, double* out) { if (!c) return -1; if (out) for (int q = 0; q < 8; q++) out[q] = 0.0; return 0; }
// the product's lincomb_kernel / wrms_kernel (vector_kernels.cuh) through the emulator
int eulerb200_vec_lincomb(eulerb200_ctx* c, int32_t nterms, const double* coef, const double* const* x, double* out,
                          int64_t n, void*)
{
  if (!c || nterms < 1 || nterms > 16) return -1;
  emu_lincomb(nterms, coef, x, out, (long)n);
  c->launches++;
  return 0;
}
int eulerb200_vec_wrms(eulerb200_ctx* c, const double* const* x, const double* const* y, double rtol, double atol,
                       int64_t nglobal, double* result, void*)
{
  const int64_t N = c->cfg.nxl * c->cfg.nyl * c->cfg.nzl;
  double acc = 0.0;
  for (int f = 0; f < 5 + (c->cfg.nchem > 0 ? 1 : 0); f++) {
    emu_wrms_accum(x[f], y[f], rtol, atol, (long)(f < 5 ? N : N * c->cfg.nchem), &acc);
    c->launches++;
  }
  *result = std::sqrt(acc / (double)nglobal);
  return 0;
}
// multi-rank entry points the drop-in references but never reaches with one rank
int eulerb200_comm_unique_id(void*) { return -1; }
int eulerb200_comm_attach(eulerb200_ctx*, const void*) { return -1; }
int eulerb200_p2p_export(eulerb200_ctx*, void*) { return -1; }
int eulerb200_p2p_attach(eulerb200_ctx*, const void*) { return -1; }
}
