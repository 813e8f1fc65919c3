// comm.cu — multi-GPU plumbing: one process per GPU, particles sharded on the
// particle index (rank r holds the contiguous slice [r*N/g, (r+1)*N/g) of every
// uncertain variable).  Elementwise work needs no communication; only reduction
// partials (six moments, M means, one M x M Gram block) cross NVLink, through
// NCCL.  The reference has no distributed path at all (`bypmap` is a plain map,
// src/bymap.jl:90-108).
//
// NCCL is dlopen'ed at first use so that libmcm_b200.so has no link-time
// dependency; a host program that already loaded NCCL (torch) shares it.
#include <dlfcn.h>
#include <string.h>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_moments.cuh"

namespace mcm {

// minimal NCCL ABI (stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };      // ncclRedOp_t

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};
static Nccl g_nccl;

static int nccl_load() {
  Nccl& n = g_nccl;
  if (n.h) return MCM_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.h) break;
  }
  if (!n.h) return fail(MCM_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define MCM_SYM(field, name)                                         \
  *(void**)(&n.field) = dlsym(n.h, name);                            \
  if (!n.field) return fail(MCM_ERR_NCCL, "libnccl: missing symbol %s", name);
  MCM_SYM(GetUniqueId, "ncclGetUniqueId")
  MCM_SYM(CommInitRank, "ncclCommInitRank")
  MCM_SYM(CommDestroy, "ncclCommDestroy")
  MCM_SYM(AllReduce, "ncclAllReduce")
  MCM_SYM(AllGather, "ncclAllGather")
  MCM_SYM(GetErrorString, "ncclGetErrorString")
#undef MCM_SYM
  return MCM_OK;
}

#define MCM_NCCL(x)                                                                              \
  do {                                                                                           \
    ncclResult_t r__ = (x);                                                                      \
    if (r__ != 0) return fail(MCM_ERR_NCCL, "%s -> %s", #x, g_nccl.GetErrorString(r__));        \
  } while (0)

struct Mom6 { double v[6]; };
int moments_device(int dtype, const void* in, size_t n, double* d_out6, const MomExchange* pxp);
int window_lookup(int id, void** peers16, int* world, int* rank, size_t* bytes);
int window_drop(int id);
double* moments_result_slot();
int read_back(double* host_dst, const double* dev_src, size_t count);
int gram_dispatch(int dtype, const void* const* vars, int M, size_t n, double scale, const double* center_host,
                  double* cov_out_dev, double* means_out_dev);

// ---- internal collectives for other translation units (select_batch.cu) ----------------------------
// ncclDataType_t values: uint8 = 1, uint32 = 3, uint64 = 5
int comm_world(int* nranks, int* rank) {
  *nranks = g_nccl.comm ? g_nccl.nranks : 1;
  *rank = g_nccl.comm ? g_nccl.rank : 0;
  return MCM_OK;
}
int comm_allreduce_sum_u32(uint32_t* d, size_t count) {
  if (!g_nccl.comm) return MCM_OK;
  MCM_NCCL(g_nccl.AllReduce(d, d, count, 3, ncclSum, g_nccl.comm, ctx().stream));
  return MCM_OK;
}
int comm_allreduce_sum_u64(unsigned long long* d, size_t count) {
  if (!g_nccl.comm) return MCM_OK;
  MCM_NCCL(g_nccl.AllReduce(d, d, count, 5, ncclSum, g_nccl.comm, ctx().stream));
  return MCM_OK;
}
int comm_allgather_bytes(const void* send, void* recv, size_t bytes_per_rank) {
  if (!g_nccl.comm) {
    if (send != recv) MCM_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx().stream));
    return MCM_OK;
  }
  MCM_NCCL(g_nccl.AllGather(send, recv, bytes_per_rank, 1, g_nccl.comm, ctx().stream));
  return MCM_OK;
}

// ---- the in-kernel moment exchange (mcm_moments.cuh: mom_exchange) ------------------------------------------------
// One 4 KiB peer window per rank, created on first use (COLLECTIVE, like every sharded entry point) and zeroed
// before any rank can store into it.  Every sharded reduction takes the next epoch; all ranks must issue their
// sharded reductions in the same order (the usual collective contract).
static int g_mx_window = -1;
static void* g_mx_peers[16] = {};
static unsigned long long g_mx_epoch = 0;
extern "C" int mcm_window_create(size_t bytes, int* id_out, void** local_out);
extern "C" int mcm_comm_fence(void);

int mom_exchange_next(MomExchange* px) {
  memset(px, 0, sizeof(*px));
  px->world = 1;
  px->flags = ctx().d_flags;
  if (!g_nccl.comm || g_nccl.nranks == 1) return MCM_OK;
  if (g_nccl.nranks > MX_MAXRANKS) return fail(MCM_ERR_UNSUPPORTED, "in-kernel moment exchange: at most %d ranks", MX_MAXRANKS);
  if (g_mx_window < 0) {
    const size_t bytes = (size_t)MX_SLOTS * MX_MAXRANKS * MX_REC * sizeof(double);
    void* local = nullptr;
    int id = -1;
    int rc = mcm_window_create(bytes, &id, &local);
    if (rc) return rc;
    MCM_CUDA(cudaMemsetAsync(local, 0, bytes, ctx().stream));
    rc = mcm_comm_fence();                      // nobody stores into a window that is still being zeroed
    if (rc) return rc;
    MCM_CUDA(cudaStreamSynchronize(ctx().stream));
    int world, rank; size_t b;
    rc = window_lookup(id, g_mx_peers, &world, &rank, &b);
    if (rc) return rc;
    g_mx_window = id;
    g_mx_epoch = 0;
  }
  for (int r = 0; r < g_nccl.nranks; ++r) px->peer[r] = static_cast<double*>(g_mx_peers[r]);
  px->world = g_nccl.nranks;
  px->rank = g_nccl.rank;
  px->epoch = ++g_mx_epoch;
  return MCM_OK;
}

__global__ void k_scale(double* a, size_t n, double s) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] *= s;
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_moments_merge(const double* a, const double* b, double* out);
extern "C" int mcm_colmeans(int dtype, const void* const* vars, int M, size_t n, double* means_out_host);

extern "C" int mcm_comm_unique_id(void* id128) {
  if (!id128) return fail(MCM_ERR_INVALID, "mcm_comm_unique_id: null pointer");
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  MCM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return MCM_OK;
}

extern "C" int mcm_comm_init(const void* id128, int nranks, int rank) {
  MCM_NEED_INIT();
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(MCM_ERR_INVALID, "mcm_comm_init: bad arguments");
  int rc = nccl_load();
  if (rc) return rc;
  if (g_nccl.comm) return fail(MCM_ERR_INVALID, "mcm_comm_init: communicator already exists");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  MCM_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
  g_nccl.nranks = nranks;
  g_nccl.rank = rank;
  return MCM_OK;
}

extern "C" int mcm_comm_destroy(void) {
  if (g_mx_window >= 0) { window_drop(g_mx_window); g_mx_window = -1; }
  if (g_nccl.comm) {
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
    g_nccl.nranks = 1;
    g_nccl.rank = 0;
  }
  return MCM_OK;
}

extern "C" int mcm_comm_info(int* nranks, int* rank) {
  if (nranks) *nranks = g_nccl.nranks;
  if (rank) *rank = g_nccl.rank;
  return MCM_OK;
}

extern "C" int mcm_comm_allreduce_sum_f64(double* dptr, size_t count) {
  MCM_NEED_INIT();
  if (!g_nccl.comm) {
    if (g_nccl.nranks == 1) return MCM_OK;  // single rank: identity
    return fail(MCM_ERR_NCCL, "mcm_comm_allreduce_sum_f64: no communicator");
  }
  MCM_NCCL(g_nccl.AllReduce(dptr, dptr, count, ncclFloat64, ncclSum, g_nccl.comm, ctx().stream));
  return MCM_OK;
}

// All ranks: one pass over the local shard; the last CTA stores the rank's partial into every rank's peer window
// and merges all partials in rank order (mom_exchange) — every rank returns bit-identical results after ONE
// device->host read.  (First version: all-gather of 6 doubles per rank + host merge.)
extern "C" int mcm_moments_sharded(int dtype, const void* in, size_t n_local, double* out_host) {
  MCM_NEED_INIT();
  if (!out_host) return fail(MCM_ERR_INVALID, "mcm_moments_sharded: null output");
  Ctx& c = ctx();
  double* slot = moments_result_slot();      // 6 doubles
  MomExchange px;
  int rc = mom_exchange_next(&px);
  if (rc) return rc;
  // an empty shard still takes part in the exchange: reduce a zero-length range (the kernel handles n == 0)
  static double* d_dummy = nullptr;
  if (!d_dummy) MCM_CUDA(cudaMalloc((void**)&d_dummy, 16));
  rc = moments_device(dtype, n_local ? in : d_dummy, n_local, slot, &px);
  if (rc) return rc;
  (void)c;
  return read_back(out_host, slot, 6);
}

// Merge per-rank moments that already sit on the host (e.g. the fused statistics of mcm_chain):
// upload 6 doubles, all-gather, Chan-merge in rank order.  In place.
extern "C" int mcm_moments_allmerge(double* inout_host6) {
  MCM_NEED_INIT();
  if (!inout_host6) return fail(MCM_ERR_INVALID, "mcm_moments_allmerge: null pointer");
  const int g = g_nccl.comm ? g_nccl.nranks : 1;
  if (g == 1) return MCM_OK;
  if (g > 64) return fail(MCM_ERR_UNSUPPORTED, "mcm_moments_allmerge: at most 64 ranks");
  Ctx& c = ctx();
  double* slot = moments_result_slot();
  double* gathered = c.d_scratch + (c.scratch_doubles - 16 - 6 * 64);
  MCM_CUDA(cudaMemcpyAsync(slot, inout_host6, 6 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  MCM_NCCL(g_nccl.AllGather(slot, gathered, 6, ncclFloat64, g_nccl.comm, c.stream));
  std::vector<double> h(6 * (size_t)g);
  int rc = read_back(h.data(), gathered, 6 * (size_t)g);
  if (rc) return rc;
  double acc[6] = {h[0], h[1], h[2], h[3], h[4], h[5]};
  for (int r = 1; r < g; ++r) {
    double t[6];
    mcm_moments_merge(acc, &h[6 * r], t);
    memcpy(acc, t, sizeof(t));
  }
  memcpy(inout_host6, acc, sizeof(acc));
  return MCM_OK;
}

// Sharded pcov: (1) global column means = allreduce(local sums)/N; (2) local centred Gram on
// the global means (raw, unscaled); (3) allreduce(sum) of the M x M block; (4) scale.
extern "C" int mcm_gram_sharded(int dtype, const void* const* vars, int M, size_t n_local, int corrected,
                                double* cov_out_dev) {
  MCM_NEED_INIT();
  if (!vars || !cov_out_dev || M < 1) return fail(MCM_ERR_INVALID, "mcm_gram_sharded: bad arguments");
  Ctx& c = ctx();
  const int g = g_nccl.comm ? g_nccl.nranks : 1;
  std::vector<double> means((size_t)M + 1);
  int rc = MCM_OK;
  if (n_local) rc = mcm_colmeans(dtype, vars, M, n_local, means.data());
  if (rc) return rc;
  // local sums + local count -> allreduce
  for (int j = 0; j < M; ++j) means[j] *= (double)n_local;
  means[M] = (double)n_local;
  double* d_tmp = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_tmp, ((size_t)M + 1) * sizeof(double), c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_tmp, means.data(), ((size_t)M + 1) * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  if (g > 1) MCM_NCCL(g_nccl.AllReduce(d_tmp, d_tmp, (size_t)M + 1, ncclFloat64, ncclSum, g_nccl.comm, c.stream));
  MCM_CUDA(cudaMemcpyAsync(means.data(), d_tmp, ((size_t)M + 1) * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d_tmp, c.stream));
  const double n_glob = means[M];
  for (int j = 0; j < M; ++j) means[j] /= n_glob;
  rc = gram_dispatch(dtype, vars, M, n_local, 1.0, means.data(), cov_out_dev, nullptr);
  if (rc) return rc;
  if (g > 1)
    MCM_NCCL(g_nccl.AllReduce(cov_out_dev, cov_out_dev, (size_t)M * M, ncclFloat64, ncclSum, g_nccl.comm, c.stream));
  const double scale = 1.0 / (n_glob - (corrected ? 1.0 : 0.0));
  k_scale<<<c.sm_count * 4, 256, 0, c.stream>>>(cov_out_dev, (size_t)M * M, scale);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
