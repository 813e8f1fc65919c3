// api.cu — lifecycle, storage and error plumbing of the C ABI (include/mcm_b200.h).
// Storage replaces the `particles::Vector{T}` field (src/types.jl:17-19): every
// reference primitive allocates a fresh result vector
// (src/register_primitive.jl:45,49,53,95); here that is a stream-ordered
// cudaMallocAsync from a pool that never returns memory to the driver, so an
// alloc/free pair costs ~1 us and never synchronises.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "mcm_common.cuh"

namespace mcm {

static Ctx g_ctx;
static thread_local char g_err[512] = "";
static std::mutex g_mu;

Ctx& ctx() { return g_ctx; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_version(void) { return MCM_B200_VERSION; }

extern "C" const char* mcm_last_error(void) { return g_err; }

extern "C" int mcm_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  Ctx& c = g_ctx;
  if (c.inited) {
    if (c.device != device) return fail(MCM_ERR_INVALID, "mcm_init: already bound to device %d", c.device);
    return MCM_OK;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(MCM_ERR_CUDA, "mcm_init: no CUDA device (%s); this engine has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(MCM_ERR_INVALID, "mcm_init: device %d out of range [0,%d)", device, ndev);
  MCM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MCM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(MCM_ERR_UNSUPPORTED, "mcm_init: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  MCM_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
  c.stream = c.own_stream;
  // keep freed blocks in the pool (no trimming at sync points)
  cudaMemPool_t pool;
  MCM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  MCM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  MCM_CUDA(cudaMalloc(&c.d_flags, sizeof(uint32_t)));
  MCM_CUDA(cudaMemset(c.d_flags, 0, sizeof(uint32_t)));
  MCM_CUDA(cudaMalloc(&c.d_tickets, kMaxTickets * sizeof(unsigned int)));
  MCM_CUDA(cudaMemset(c.d_tickets, 0, kMaxTickets * sizeof(unsigned int)));
  c.scratch_doubles = kScratchDoubles;
  MCM_CUDA(cudaMalloc(&c.d_scratch, c.scratch_doubles * sizeof(double)));
  c.pinned_bytes = kPinnedBytes;
  MCM_CUDA(cudaMallocHost(&c.h_pinned, c.pinned_bytes));
  c.inited = true;
  return MCM_OK;
}

extern "C" int mcm_comm_destroy(void);

extern "C" int mcm_finalize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  Ctx& c = g_ctx;
  if (!c.inited) return MCM_OK;
  mcm_comm_destroy();
  cudaStreamSynchronize(c.stream);
  cudaFree(c.d_flags);
  cudaFree(c.d_tickets);
  cudaFree(c.d_scratch);
  cudaFreeHost(c.h_pinned);
  cudaStreamDestroy(c.own_stream);
  c = Ctx();
  return MCM_OK;
}

extern "C" int mcm_set_stream(void* cuda_stream) {
  MCM_NEED_INIT();
  Ctx& c = ctx();
  // order the new stream after everything already queued on the old one
  cudaStream_t ns = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c.own_stream;
  if (ns != c.stream) {
    cudaEvent_t ev;
    MCM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MCM_CUDA(cudaEventRecord(ev, c.stream));
    MCM_CUDA(cudaStreamWaitEvent(ns, ev, 0));
    MCM_CUDA(cudaEventDestroy(ev));
    c.stream = ns;
  }
  return MCM_OK;
}

extern "C" void* mcm_get_stream(void) { return ctx().inited ? (void*)ctx().stream : nullptr; }

extern "C" int mcm_sync(void) {
  MCM_NEED_INIT();
  MCM_CUDA(cudaStreamSynchronize(ctx().stream));
  return MCM_OK;
}

// clear only the sticky bits in `mask` (stream-ordered); the others keep waiting for their own check
extern "C" int mcm_clear_flags(uint32_t mask) {
  MCM_NEED_INIT();
  Ctx& c = ctx();
  MCM_CUDA(cudaMemcpyAsync(c.h_pinned, c.d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  uint32_t* h = static_cast<uint32_t*>(c.h_pinned);
  *h &= ~mask;
  MCM_CUDA(cudaMemcpyAsync(c.d_flags, h, sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  return MCM_OK;
}

extern "C" int mcm_check_flags(uint32_t* bits, int reset) {
  MCM_NEED_INIT();
  Ctx& c = ctx();
  if (!bits) return fail(MCM_ERR_INVALID, "mcm_check_flags: null pointer");
  MCM_CUDA(cudaMemcpyAsync(c.h_pinned, c.d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  if (reset) MCM_CUDA(cudaMemsetAsync(c.d_flags, 0, sizeof(uint32_t), c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  *bits = *static_cast<uint32_t*>(c.h_pinned);
  return MCM_OK;
}

extern "C" int mcm_device_info(int* sm_count, size_t* free_bytes, size_t* total_bytes) {
  MCM_NEED_INIT();
  if (sm_count) *sm_count = ctx().sm_count;
  size_t f = 0, t = 0;
  MCM_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return MCM_OK;
}

extern "C" int mcm_alloc(void** dptr, size_t bytes) {
  MCM_NEED_INIT();
  if (!dptr) return fail(MCM_ERR_INVALID, "mcm_alloc: null pointer");
  *dptr = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocAsync(dptr, bytes, ctx().stream);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    return fail(MCM_ERR_NOMEM, "mcm_alloc: out of device memory (%zu bytes)", bytes);
  }
  MCM_CUDA(e);
  return MCM_OK;
}

namespace mcm {
void* pinned_scratch(int slot, size_t bytes) {
  static void* buf[8] = {};
  static size_t cap[8] = {};
  if (slot < 0 || slot >= 8) return nullptr;
  if (bytes > cap[slot]) {
    if (buf[slot]) cudaFreeHost(buf[slot]);
    buf[slot] = nullptr; cap[slot] = 0;
    const size_t want = bytes + bytes / 4;
    if (cudaHostAlloc(&buf[slot], want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); buf[slot] = nullptr; return nullptr; }
    cap[slot] = want;
  }
  return buf[slot];
}
}  // namespace mcm

// M result vectors of one call (a vector of uncertain variables): M pool allocations in one C call — through ctypes
// each costs ~5 us of interpreter time, which at M = 1024 is as long as the gather that fills them
extern "C" int mcm_alloc_many(void** dptrs, int M, size_t bytes_each) {
  MCM_NEED_INIT();
  if (!dptrs || M < 0) return fail(MCM_ERR_INVALID, "mcm_alloc_many: bad arguments");
  if (bytes_each == 0) bytes_each = 16;
  cudaStream_t st = ctx().stream;
  for (int j = 0; j < M; ++j) {
    dptrs[j] = nullptr;
    const cudaError_t e = cudaMallocAsync(&dptrs[j], bytes_each, st);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int k = 0; k < j; ++k) { cudaFreeAsync(dptrs[k], st); dptrs[k] = nullptr; }
      if (e == cudaErrorMemoryAllocation) return fail(MCM_ERR_NOMEM, "mcm_alloc_many: out of device memory (%d x %zu bytes)", M, bytes_each);
      MCM_CUDA(e);
    }
  }
  return MCM_OK;
}

extern "C" int mcm_free(void* dptr) {
  if (!dptr) return MCM_OK;
  if (!ctx().inited) return MCM_OK;  // finalizers may run after mcm_finalize
  MCM_CUDA(cudaFreeAsync(dptr, ctx().stream));
  return MCM_OK;
}

extern "C" int mcm_host_alloc(void** hptr, size_t bytes) {
  MCM_NEED_INIT();
  if (!hptr) return fail(MCM_ERR_INVALID, "mcm_host_alloc: null pointer");
  MCM_CUDA(cudaMallocHost(hptr, bytes ? bytes : 16));
  return MCM_OK;
}

// write-combined pinned memory: for buffers the host only WRITES and the device reads (inputs of mcm_chain_host):
// the DMA reads need no cache snoop.  Host reads of such memory are very slow.
extern "C" int mcm_host_alloc_wc(void** hptr, size_t bytes) {
  MCM_NEED_INIT();
  if (!hptr) return fail(MCM_ERR_INVALID, "mcm_host_alloc_wc: null pointer");
  MCM_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocWriteCombined));
  return MCM_OK;
}

extern "C" int mcm_host_free(void* hptr) {
  if (!hptr) return MCM_OK;
  MCM_CUDA(cudaFreeHost(hptr));
  return MCM_OK;
}

extern "C" int mcm_upload(void* dst_dev, const void* src_host, size_t bytes) {
  MCM_NEED_INIT();
  if (bytes == 0) return MCM_OK;
  if (!dst_dev || !src_host) return fail(MCM_ERR_INVALID, "mcm_upload: null pointer");
  // pageable sources are staged by the runtime before the call returns; pinned
  // sources are truly asynchronous (caller keeps them alive until mcm_sync)
  MCM_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx().stream));
  return MCM_OK;
}

extern "C" int mcm_download(void* dst_host, const void* src_dev, size_t bytes) {
  MCM_NEED_INIT();
  if (bytes == 0) return MCM_OK;
  if (!dst_host || !src_dev) return fail(MCM_ERR_INVALID, "mcm_download: null pointer");
  MCM_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  MCM_CUDA(cudaStreamSynchronize(ctx().stream));
  return MCM_OK;
}

extern "C" int mcm_copy(void* dst_dev, const void* src_dev, size_t bytes) {
  MCM_NEED_INIT();
  if (bytes == 0) return MCM_OK;
  if (!dst_dev || !src_dev) return fail(MCM_ERR_INVALID, "mcm_copy: null pointer");
  MCM_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx().stream));
  return MCM_OK;
}
