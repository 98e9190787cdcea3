/* ---------------------------------------------------------------------------
 * shim_core.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A minimal, in-process stand-in for the handful of SUNDIALS 6.2 and MPI symbols
 * that the reference's fluid-RHS path touches, so that
 *     /root/reference/src/utilities.cpp  (fEuler, face_flux, stability, check_flag)
 *     /root/reference/src/euler3D.hpp    (class EulerData, halo exchange, pack1D_*)
 *     /root/reference/src/profiler.hpp
 * compile UNMODIFIED (from where they lie) into oracle/_ref/ with plain g++.
 * Nothing in here is copied from SUNDIALS or from an MPI implementation: the
 * types are the smallest things that satisfy the call sites
 * (euler3D.hpp:17-42 lists the includes; SURVEY.md section 8(c) lists the symbols).
 *
 * "MPI" here is a set of virtual ranks, each a std::thread of the same process,
 * exchanging messages through a mutex-protected mailbox.  With one virtual rank a
 * periodic self-send is matched to the self-receive by tag, which is exactly how
 * the reference's periodic wrap works on one rank.
 * ------------------------------------------------------------------------- */
#ifndef EULERB200_ORACLE_SHIM_CORE_H
#define EULERB200_ORACLE_SHIM_CORE_H

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <stdint.h>
#include <vector>

/* ------------------------------- SUNDIALS ------------------------------ */
typedef double realtype;
typedef long int sunindextype;
typedef int booleantype;
#define RCONST(x) x
#define SUNDIALS_DOUBLE_PRECISION 1
/* SUNDIALS 6.2 sundials_math.h semantics: non-positive argument gives zero. */
#define SUNRsqrt(x) ((x) <= RCONST(0.0) ? (RCONST(0.0)) : (std::sqrt((x))))
#define SUNRabs(x) (std::fabs((x)))

struct shim_SUNContext_ { int dummy; };
typedef shim_SUNContext_* SUNContext;
static inline int SUNContext_Create(void* /*comm*/, SUNContext* ctx) {
  *ctx = new shim_SUNContext_();
  return 0;
}
static inline int SUNContext_Free(SUNContext* ctx) {
  if (ctx && *ctx) { delete *ctx; *ctx = NULL; }
  return 0;
}
struct shim_SUNMemoryHelper_ { int dummy; };
typedef shim_SUNMemoryHelper_* SUNMemoryHelper;
static inline int SUNMemoryHelper_Destroy(SUNMemoryHelper h) { delete h; return 0; }

/* N_Vector: either a leaf (owns or borrows one contiguous array) or a
 * many-vector (array of leaf pointers).  Only what fEuler/ExchangeStart use. */
struct shim_NVector_ {
  sunindextype length;
  realtype* data;
  int own;             /* 0 borrowed, 1 new[], 2 managed memory (SHIM_MANAGED_VECTORS) */
  int nsub;
  shim_NVector_** sub;
};
typedef shim_NVector_* N_Vector;

/* -DSHIM_MANAGED_VECTORS (the refmain_gpu_* programs): the leaf vectors the driver creates with
 * N_VNew_Serial live in CUDA managed memory, like the N_VNewManaged_* vectors of the reference's own
 * device builds -- the unmodified driver's host code keeps working on them, the drop-in fEuler and
 * the vector operations of shim_arkstep.cpp run on the device with the state resident there. */
#ifdef SHIM_MANAGED_VECTORS
extern "C" void* eulerb200_managed_alloc(int64_t bytes);
extern "C" void eulerb200_device_free(void* p);
#endif

static inline N_Vector N_VNew_Serial(sunindextype n, SUNContext) {
  N_Vector v = new shim_NVector_();
#ifdef SHIM_MANAGED_VECTORS
  v->length = n; v->data = (realtype*)eulerb200_managed_alloc((int64_t)sizeof(realtype) * n); v->own = 2;
  if (!v->data) { fprintf(stderr, "shim: eulerb200_managed_alloc failed\n"); abort(); }
  memset(v->data, 0, sizeof(realtype) * n);
  v->nsub = 0; v->sub = NULL;
#else
  v->length = n; v->data = new realtype[n](); v->own = 1; v->nsub = 0; v->sub = NULL;
#endif
  return v;
}
static inline N_Vector N_VMake_Serial(sunindextype n, realtype* data, SUNContext) {
  N_Vector v = new shim_NVector_();
  v->length = n; v->data = data; v->own = 0; v->nsub = 0; v->sub = NULL;
  return v;
}
static inline realtype* N_VGetArrayPointer(N_Vector v) { return v ? v->data : NULL; }
static inline void N_VDestroy(N_Vector v) {
  if (!v) return;
#ifdef SHIM_MANAGED_VECTORS
  if (v->own == 2 && v->data) eulerb200_device_free(v->data);
#endif
  if (v->own == 1 && v->data) delete[] v->data;
  if (v->sub) delete[] v->sub;
  delete v;
}
static inline void N_VConst(realtype c, N_Vector v) {
  if (v->nsub > 0) { for (int s = 0; s < v->nsub; s++) N_VConst(c, v->sub[s]); return; }
  for (sunindextype i = 0; i < v->length; i++) v->data[i] = c;
}

/* --------------------------------- MPI --------------------------------- */
struct shim_Comm_ {
  int cart;         /* 0: world, 1: cartesian */
  int dims[3];
  int periods[3];
};
typedef shim_Comm_* MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
struct MPI_Request_ { int kind; void* buf; int count; int peer; int tag; int done; };
typedef MPI_Request_* MPI_Request;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_SUCCESS 0
#define MPI_PROC_NULL (-2)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_DOUBLE 1
#define MPI_LONG 2
#define MPI_INT 3
#define MPI_BYTE 4
#define MPI_SUNREALTYPE MPI_DOUBLE
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_IN_PLACE ((void*)-1)

extern shim_Comm_ shim_world_comm;
#define MPI_COMM_WORLD (&shim_world_comm)

/* implemented in shim_mpi.cpp */
void shim_set_world(int nprocs);            /* before launching rank threads */
void shim_set_rank(int rank);               /* first thing in each rank thread */
int MPI_Comm_size(MPI_Comm, int* size);
int MPI_Comm_rank(MPI_Comm, int* rank);
int MPI_Dims_create(int nnodes, int ndims, int* dims);
int MPI_Cart_create(MPI_Comm, int ndims, const int* dims, const int* periods, int reorder, MPI_Comm* out);
int MPI_Cart_get(MPI_Comm, int maxdims, int* dims, int* periods, int* coords);
int MPI_Cart_rank(MPI_Comm, const int* coords, int* rank);
int MPI_Irecv(void* buf, int count, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Request* req);
int MPI_Isend(const void* buf, int count, MPI_Datatype, int dst, int tag, MPI_Comm, MPI_Request* req);
int MPI_Waitall(int n, MPI_Request* req, MPI_Status* stat);
int MPI_Allreduce(const void* in, void* out, int count, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype, int root, MPI_Comm);
int MPI_Allgather(const void* in, int incount, MPI_Datatype, void* out, int outcount, MPI_Datatype, MPI_Comm);
int MPI_Abort(MPI_Comm, int code);
double MPI_Wtime(void);
/* what the reference's main programs need on top (euler3D_main.cpp, io.cpp, problem files) */
static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Reduce(const void* in, void* out, int count, MPI_Datatype t, MPI_Op op, int /*root*/, MPI_Comm c) {
  return MPI_Allreduce(in, out, count, t, op, c);     /* every rank gets the result: a superset */
}

static inline N_Vector N_VMake_MPIManyVector(MPI_Comm, sunindextype nsub, N_Vector* subs, SUNContext) {
  N_Vector v = new shim_NVector_();
  v->length = 0; v->data = NULL; v->own = 0; v->nsub = (int)nsub;
  v->sub = new shim_NVector_*[nsub];
  for (int s = 0; s < (int)nsub; s++) { v->sub[s] = subs[s]; v->length += subs[s]->length; }
  return v;
}
static inline realtype* N_VGetSubvectorArrayPointer_MPIManyVector(N_Vector v, sunindextype i) {
  if (!v || i < 0 || i >= v->nsub) return NULL;
  return v->sub[i]->data;
}

static inline N_Vector N_VGetSubvector_MPIManyVector(N_Vector v, sunindextype i) {
  return (!v || i < 0 || i >= v->nsub) ? NULL : v->sub[i];
}
static inline void N_VScale(realtype c, N_Vector x, N_Vector z) {
  if (x->nsub > 0) { for (int s = 0; s < x->nsub; s++) N_VScale(c, x->sub[s], z->sub[s]); return; }
  for (sunindextype i = 0; i < x->length; i++) z->data[i] = c * x->data[i];
}
static inline int N_VEnableFusedOps_Serial(N_Vector, booleantype) { return 0; }
static inline int N_VEnableFusedOps_MPIManyVector(N_Vector, booleantype) { return 0; }

/* ------------------------------- ARKODE -------------------------------- */
/* Only the enum *types* are needed (class ARKODEParameters, euler3D.hpp:126-172). */
typedef enum { ARKODE_ERK_NONE = -1 } ARKODE_ERKTableID;
typedef enum { ARKODE_DIRK_NONE = -1 } ARKODE_DIRKTableID;
typedef enum { ARKODE_MRI_NONE = -1 } ARKODE_MRITableID;

#endif
