/* ---------------------------------------------------------------------------
 * shim_mpi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Virtual-rank "MPI": every rank is a std::thread of this process; point-to-point
 * messages are copied eagerly into a mailbox keyed by (src, dst, tag) at Isend
 * time and picked up in Waitall.  Only the calls the reference's fluid-RHS path
 * makes are provided.
 * ------------------------------------------------------------------------- */
#include "shim_core.h"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <tuple>
#include <deque>

shim_Comm_ shim_world_comm = {0, {1, 1, 1}, {0, 0, 0}};

namespace {
int g_nprocs = 1;
thread_local int t_rank = 0;

std::mutex g_mtx;
std::condition_variable g_cv;
typedef std::tuple<int, int, int> Key;   /* src, dst, tag */
std::map<Key, std::deque<std::vector<double> > > g_mail;

/* sense-reversing barrier + reduction scratch */
int g_bar_count = 0;
int g_bar_gen = 0;
std::vector<double> g_red;
int g_red_count = 0;
}  // namespace

void shim_set_world(int nprocs) {
  std::lock_guard<std::mutex> lk(g_mtx);
  g_nprocs = nprocs;
  g_mail.clear();
  g_bar_count = 0;
  g_red_count = 0;
}
void shim_set_rank(int rank) { t_rank = rank; }

int MPI_Comm_size(MPI_Comm, int* size) { *size = g_nprocs; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = t_rank; return MPI_SUCCESS; }

/* Balanced factorisation in non-increasing order (the behaviour every mainstream
 * MPI gives for the process counts used here: 1,2,4,8 -> 1x1x1, 2x1x1, 2x2x1, 2x2x2). */
int MPI_Dims_create(int nnodes, int ndims, int* dims) {
  std::vector<int> primes;
  int n = nnodes;
  for (int p = 2; p * p <= n; p++) while (n % p == 0) { primes.push_back(p); n /= p; }
  if (n > 1) primes.push_back(n);
  std::sort(primes.begin(), primes.end(), [](int a, int b) { return a > b; });
  std::vector<int> bins(ndims > 0 ? ndims : 0, 1);
  for (size_t i = 0; i < primes.size() && ndims > 0; i++) {
    int best = 0;
    for (int d = 1; d < ndims; d++) if (bins[d] < bins[best]) best = d;
    bins[best] *= primes[i];
  }
  std::sort(bins.begin(), bins.end(), [](int a, int b) { return a > b; });
  for (int d = 0; d < ndims; d++) dims[d] = bins[d];
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm, int ndims, const int* dims, const int* periods, int, MPI_Comm* out) {
  shim_Comm_* c = new shim_Comm_();   /* leaked on purpose: lives as long as the EulerData */
  c->cart = 1;
  for (int d = 0; d < 3; d++) { c->dims[d] = d < ndims ? dims[d] : 1; c->periods[d] = d < ndims ? periods[d] : 0; }
  *out = c;
  return MPI_SUCCESS;
}
/* reorder = 0 -> row-major rank order, coords[0] slowest */
int MPI_Cart_get(MPI_Comm c, int maxdims, int* dims, int* periods, int* coords) {
  int r = t_rank;
  int co[3];
  co[2] = r % c->dims[2]; r /= c->dims[2];
  co[1] = r % c->dims[1]; r /= c->dims[1];
  co[0] = r;
  for (int d = 0; d < maxdims && d < 3; d++) { dims[d] = c->dims[d]; periods[d] = c->periods[d]; coords[d] = co[d]; }
  return MPI_SUCCESS;
}
int MPI_Cart_rank(MPI_Comm c, const int* coords, int* rank) {
  int co[3];
  for (int d = 0; d < 3; d++) {
    co[d] = coords[d];
    if (c->periods[d]) co[d] = ((co[d] % c->dims[d]) + c->dims[d]) % c->dims[d];
    else if (co[d] < 0 || co[d] >= c->dims[d]) { *rank = MPI_PROC_NULL; return 1; }
  }
  *rank = (co[0] * c->dims[1] + co[1]) * c->dims[2] + co[2];
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Request* req) {
  MPI_Request r = new MPI_Request_();
  r->kind = 0; r->buf = buf; r->count = count; r->peer = src; r->tag = tag; r->done = 0;
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype, int dst, int tag, MPI_Comm, MPI_Request* req) {
  {
    std::lock_guard<std::mutex> lk(g_mtx);
    const double* d = (const double*)buf;
    g_mail[Key(t_rank, dst, tag)].emplace_back(d, d + count);
  }
  g_cv.notify_all();
  MPI_Request r = new MPI_Request_();
  r->kind = 1; r->buf = NULL; r->count = count; r->peer = dst; r->tag = tag; r->done = 1;
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request* req, MPI_Status*) {
  for (int i = 0; i < n; i++) {
    MPI_Request r = req[i];
    if (r == MPI_REQUEST_NULL) continue;
    if (r->kind == 0) {
      std::unique_lock<std::mutex> lk(g_mtx);
      Key key(r->peer, t_rank, r->tag);
      g_cv.wait(lk, [&] { auto it = g_mail.find(key); return it != g_mail.end() && !it->second.empty(); });
      std::vector<double>& m = g_mail[key].front();
      if ((int)m.size() != r->count) { fprintf(stderr, "shim MPI: message size mismatch\n"); abort(); }
      memcpy(r->buf, m.data(), sizeof(double) * m.size());
      g_mail[key].pop_front();
    }
    delete r;
    req[i] = MPI_REQUEST_NULL;
  }
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm) {
  std::unique_lock<std::mutex> lk(g_mtx);
  int gen = g_bar_gen;
  if (++g_bar_count == g_nprocs) { g_bar_count = 0; g_bar_gen++; g_cv.notify_all(); }
  else g_cv.wait(lk, [&] { return g_bar_gen != gen; });
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* in, void* out, int count, MPI_Datatype, MPI_Op op, MPI_Comm comm) {
  const double* src = (in == MPI_IN_PLACE) ? (const double*)out : (const double*)in;
  std::vector<double> mine(src, src + count);
  {
    std::lock_guard<std::mutex> lk(g_mtx);
    if (g_red_count == 0) g_red = mine;
    else for (int i = 0; i < count; i++) {
      if (op == MPI_SUM) g_red[i] += mine[i];
      else if (op == MPI_MIN) g_red[i] = std::min(g_red[i], mine[i]);
      else g_red[i] = std::max(g_red[i], mine[i]);
    }
    g_red_count++;
  }
  MPI_Barrier(comm);
  std::vector<double> res;
  { std::lock_guard<std::mutex> lk(g_mtx); res = g_red; }
  MPI_Barrier(comm);
  { std::lock_guard<std::mutex> lk(g_mtx); g_red_count = 0; }
  MPI_Barrier(comm);
  memcpy(out, res.data(), sizeof(double) * count);
  return MPI_SUCCESS;
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  static std::vector<char> box;
  const size_t bytes = (size_t)count * (type == MPI_BYTE ? 1 : (type == MPI_INT ? 4 : 8));
  if (t_rank == root) { std::lock_guard<std::mutex> lk(g_mtx); box.assign((char*)buf, (char*)buf + bytes); }
  MPI_Barrier(comm);
  if (t_rank != root) { std::lock_guard<std::mutex> lk(g_mtx); memcpy(buf, box.data(), bytes); }
  MPI_Barrier(comm);
  return MPI_SUCCESS;
}

int MPI_Allgather(const void* in, int incount, MPI_Datatype type, void* out, int, MPI_Datatype, MPI_Comm comm)
{
  static std::vector<char> box;
  const size_t bytes = (size_t)incount * (type == MPI_BYTE ? 1 : (type == MPI_INT ? 4 : 8));
  { std::lock_guard<std::mutex> lk(g_mtx); if (box.size() != bytes * g_nprocs) box.assign(bytes * g_nprocs, 0);
    memcpy(box.data() + bytes * t_rank, in, bytes); }
  MPI_Barrier(comm);
  { std::lock_guard<std::mutex> lk(g_mtx); memcpy(out, box.data(), bytes * g_nprocs); }
  MPI_Barrier(comm);
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int code) { fprintf(stderr, "shim MPI_Abort(%d)\n", code); abort(); return 0; }

double MPI_Wtime(void) {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}
