/* ---------------------------------------------------------------------------
 * shim_arkstep.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * The ARKStep calls of the reference's euler3D_main.cpp on top of host/erk_stepper.hpp (the ERK
 * loop of the native driver) with host N_Vectors: see arkode/arkode_arkstep.h.  Options left at 0
 * by the reference's input files mean "ARKODE's default", exactly as the native driver reads them.
 * ------------------------------------------------------------------------- */
#include "arkode/arkode_arkstep.h"
#include "erk_stepper.hpp"
#ifdef SHIM_MANAGED_VECTORS
#include "eulerb200.h"
extern "C" eulerb200_ctx* eulerb200_dropin_context(void* user_data);
#endif

namespace {

struct HostVec { N_Vector v; };
void swap(HostVec& a, HostVec& b) { N_Vector t = a.v; a.v = b.v; b.v = t; }

struct NVecOps {
  typedef HostVec Vec;
  ARKRhsFn fe = NULL;
  ARKExpStabFn stab = NULL;
  void* user = NULL;
  void* stab_data = NULL;
#ifdef SHIM_MANAGED_VECTORS
  // vectors in managed memory: stage combinations and the error norm run on the device
  // (eulerb200_vec_lincomb / _wrms_accum), the state never leaves it between outputs
  eulerb200_ctx* ctx() { return eulerb200_dropin_context(user); }
  double* d_acc = NULL;
  void lincomb(Vec& out, int n, const double* c, Vec* const* v)
  {
    for (int s = 0; s < out.v->nsub; s++) {
      const double* x[16];
      for (int q = 0; q < n; q++) x[q] = v[q]->v->sub[s]->data;
      if (eulerb200_vec_lincomb(ctx(), n, c, x, out.v->sub[s]->data, out.v->sub[s]->length, NULL) != 0) abort();
    }
  }
  double wrms(const Vec& x, const Vec& y, double rtol, double atol)
  {
    if (!d_acc) d_acc = (double*)eulerb200_device_alloc(sizeof(double));
    double acc[2] = {0.0, 0.0};
    eulerb200_copy_to_device(d_acc, acc, sizeof(double));
    for (int s = 0; s < x.v->nsub; s++) {
      if (eulerb200_vec_wrms_accum(ctx(), x.v->sub[s]->data, y.v->sub[s]->data, rtol, atol, x.v->sub[s]->length, d_acc, NULL) != 0) abort();
      acc[1] += (double)x.v->sub[s]->length;
    }
    eulerb200_copy_to_host(acc, d_acc, sizeof(double));
    MPI_Allreduce(MPI_IN_PLACE, acc, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    return std::sqrt(acc[0] / acc[1]);
  }
  void copy(N_Vector src, N_Vector dst)
  {
    const double one = 1.0;
    for (int s = 0; s < src->nsub; s++) {
      const double* x[1] = {src->sub[s]->data};
      if (eulerb200_vec_lincomb(ctx(), 1, &one, x, dst->sub[s]->data, src->sub[s]->length, NULL) != 0) abort();
    }
    eulerb200_synchronize(ctx());      // the driver's host code reads dst next
  }
#else
  void copy(N_Vector src, N_Vector dst) { N_VScale(1.0, src, dst); }
  void lincomb(Vec& out, int n, const double* c, Vec* const* v)
  {
    for (int s = 0; s < out.v->nsub; s++) {
      const sunindextype len = out.v->sub[s]->length;
      double* o = out.v->sub[s]->data;
      for (sunindextype i = 0; i < len; i++) {
        double acc = c[0] * v[0]->v->sub[s]->data[i];
        for (int q = 1; q < n; q++) acc = std::fma(c[q], v[q]->v->sub[s]->data[i], acc);
        o[i] = acc;
      }
    }
  }
  double wrms(const Vec& x, const Vec& y, double rtol, double atol)
  {
    double acc[2] = {0.0, 0.0};
    for (int s = 0; s < x.v->nsub; s++)
      for (sunindextype i = 0; i < x.v->sub[s]->length; i++) {
        const double q = x.v->sub[s]->data[i] / std::fma(rtol, std::fabs(y.v->sub[s]->data[i]), atol);
        acc[0] = std::fma(q, q, acc[0]);
        acc[1] += 1.0;
      }
    MPI_Allreduce(MPI_IN_PLACE, acc, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    return std::sqrt(acc[0] / acc[1]);
  }
#endif
  int rhs(double t, Vec& y, Vec& out) { return fe(t, y.v, out.v, user); }
  int stability(Vec& w, double t, double, double* dt) { return stab(w.v, t, dt, stab_data); }
};

N_Vector clone(N_Vector y, SUNContext ctx)
{
  N_Vector sub[8];
  for (int s = 0; s < y->nsub; s++) sub[s] = N_VNew_Serial(y->sub[s]->length, ctx);
  return N_VMake_MPIManyVector(MPI_COMM_WORLD, y->nsub, sub, ctx);
}

struct ArkMem {
  ErkStepper<NVecOps> S;
  SUNContext ctx;
  int order = 4, etable = -1;
  bool started = false;
  double tstop = 0;
};

}  // namespace

void* ARKStepCreate(ARKRhsFn fe, ARKRhsFn, realtype t0, N_Vector y0, SUNContext ctx)
{
  if (!fe || !y0 || y0->nsub <= 0) return NULL;
  ArkMem* m = new ArkMem();
  m->ctx = ctx;
  m->S.ops.fe = fe;
  m->S.t = t0;
  m->S.w.v = clone(y0, ctx);
  N_VScale(1.0, y0, m->S.w.v);          // (host copy: the user data, hence the device context, is not known yet)
  m->S.ytmp.v = clone(y0, ctx);
  m->S.yerr.v = clone(y0, ctx);
  for (int i = 0; i < 13; i++) m->S.k[i].v = clone(y0, ctx);
  return m;
}
void ARKStepFree(void** mem) { if (mem && *mem) { delete (ArkMem*)*mem; *mem = NULL; } }
#define M ((ArkMem*)mem)
#define DFLT(x, d) ((x) != 0 ? (x) : (d))
int ARKStepSetUserData(void* mem, void* u) { M->S.ops.user = u; return 0; }
int ARKStepSetDiagnostics(void*, FILE*) { return 0; }
int ARKStepSetOrder(void* mem, int order) { M->order = order; return 0; }
int ARKStepSetTableNum(void* mem, ARKODE_DIRKTableID, ARKODE_ERKTableID e) { M->order = 0; M->etable = (int)e; return 0; }
int ARKStepSetDenseOrder(void*, int) { return 0; }
int ARKStepSetSafetyFactor(void* mem, realtype x) { M->S.safety = DFLT(x, 0.96); return 0; }
int ARKStepSetErrorBias(void* mem, realtype x) { M->S.bias = DFLT(x, 1.5); return 0; }
int ARKStepSetMaxGrowth(void* mem, realtype x) { M->S.growth = DFLT(x, 20.0); return 0; }
int ARKStepSetAdaptivityMethod(void* mem, int, int idefault, int, realtype* p)
{
  if (!idefault && p) { M->S.k1 = DFLT(p[0], 0.58); M->S.k2 = DFLT(p[1], 0.21); M->S.k3 = DFLT(p[2], 0.1); }
  return 0;
}
int ARKStepSetMaxFirstGrowth(void* mem, realtype x) { M->S.etamx1 = DFLT(x, 1e4); return 0; }
int ARKStepSetMaxEFailGrowth(void* mem, realtype x) { M->S.etamxf = DFLT(x, 0.3); return 0; }
int ARKStepSetInitStep(void* mem, realtype x) { M->S.h0 = x; return 0; }
int ARKStepSetMinStep(void* mem, realtype x) { M->S.hmin = x; return 0; }
int ARKStepSetMaxStep(void* mem, realtype x) { M->S.hmax = x; return 0; }
int ARKStepSetMaxErrTestFails(void* mem, int x) { M->S.maxnef = DFLT(x, 7); return 0; }
int ARKStepSetMaxHnilWarns(void*, int) { return 0; }
int ARKStepSetStabilityFn(void* mem, ARKExpStabFn f, void* d) { M->S.ops.stab = f; M->S.ops.stab_data = d; return 0; }
int ARKStepSetFixedStep(void* mem, realtype h) { M->S.fixedstep = (h != 0.0); M->S.hmax = h; M->S.h = 0.0; return 0; }
int ARKStepSetMaxNumSteps(void* mem, long int x) { M->S.mxsteps = x > 0 ? (int)x : 500; return 0; }
int ARKStepSStolerances(void* mem, realtype r, realtype a) { M->S.rtol = r; M->S.atol = a; return 0; }
int ARKStepSetStopTime(void* mem, realtype t) { M->tstop = t; return 0; }
int ARKStepEvolve(void* mem, realtype tout, N_Vector yout, realtype* tret, int)
{
  if (!M->started) {
    if (!make_table(M->order, M->etable, M->S.T)) { fprintf(stderr, "shim ARKStep: no Butcher table for order %d / table %d\n", M->order, M->etable); return -1; }
    if (M->S.ops.stab) M->S.cfl = 1.0;            /* the user's function already applies its own cfl factor */
    M->started = true;
  }
  const int rc = M->S.evolve(tout);
  M->S.ops.copy(M->S.w.v, yout);
  *tret = M->S.t;
  return rc;
}
int ARKStepGetCurrentStep(void* mem, realtype* h) { *h = M->S.h; return 0; }
int ARKStepGetNumSteps(void* mem, long int* n) { *n = M->S.nst; return 0; }
int ARKStepGetNumStepAttempts(void* mem, long int* n) { *n = M->S.nst_a; return 0; }
int ARKStepGetNumRhsEvals(void* mem, long int* nfe, long int* nfi) { *nfe = M->S.nfe; *nfi = 0; return 0; }
int ARKStepGetNumErrTestFails(void* mem, long int* n) { *n = M->S.netf; return 0; }
