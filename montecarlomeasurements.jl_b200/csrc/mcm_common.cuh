// mcm_common.cuh — shared host/device plumbing for libmcm_b200.so (sm_100a only).
#pragma once
// Under NVRTC (runtime-specialised chain kernels, chain_jit.cu) only the device half of this header is compiled;
// <stdint.h> / <stddef.h> / mcm_b200.h are then the small stand-ins chain_jit.cu hands to the compiler.
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stddef.h>
#ifdef __CUDACC_RTC__
#include "mcm_b200.h"
#else
#include "../../include/mcm_b200.h"
#endif

namespace mcm {
#ifdef MCM_JIT_THREADS
static constexpr int kThreads = MCM_JIT_THREADS;   // generated chain kernels may choose their CTA size (chain_jit.cu)
#else
static constexpr int kThreads = 256;          // CTA size for streaming kernels
#endif
}

#ifndef __CUDACC_RTC__
namespace mcm {

// One context per process (one process per GPU).
struct Ctx {
  bool inited = false;
  int device = -1;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;  // created by mcm_init
  cudaStream_t stream = nullptr;      // stream all work is ordered on
  uint32_t* d_flags = nullptr;        // sticky numerical flags (device)
  unsigned int* d_tickets = nullptr;  // last-block-done tickets (kept zero between launches)
  double* d_scratch = nullptr;        // reduction partials / small device workspaces
  size_t scratch_doubles = 0;
  void* h_pinned = nullptr;           // pinned readback buffer
  size_t pinned_bytes = 0;
};

Ctx& ctx();
int fail(int code, const char* fmt, ...);
// grow-only pinned host staging buffers (slots 0..7), for host <-> device copies of MB-sized tables: pageable copies run
// at a few GB/s on this class of VM and synchronise; returns nullptr if the allocation fails
void* pinned_scratch(int slot, size_t bytes);

static constexpr size_t kScratchDoubles = 1u << 17;  // 1 MiB of doubles
static constexpr size_t kPinnedBytes = 1u << 20;
static constexpr int kMaxTickets = 64;

// Grid for a grid-stride streaming kernel: `work` = number of per-thread work
// items; never more than ctas_per_sm full waves of the 148 SMs (persistent).
inline int stream_grid(size_t work, int items_per_thread, int ctas_per_sm) {
  size_t per_cta = (size_t)kThreads * (size_t)items_per_thread;
  size_t want = (work + per_cta - 1) / per_cta;
  size_t cap = (size_t)ctx().sm_count * (size_t)ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace mcm

#define MCM_CUDA(x)                                                                      \
  do {                                                                                   \
    cudaError_t e__ = (x);                                                               \
    if (e__ != cudaSuccess)                                                              \
      return mcm::fail(MCM_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #x,          \
                       cudaGetErrorString(e__));                                         \
  } while (0)

#define MCM_NEED_INIT()                                                                  \
  do {                                                                                   \
    if (!mcm::ctx().inited) return mcm::fail(MCM_ERR_NOT_INIT, "mcm_init was not called"); \
  } while (0)

#define MCM_CHECK_LAUNCH() MCM_CUDA(cudaGetLastError())

namespace mcm {
// Row statistics (digit-split exponents and max/sigma ratios about the column means) handed from the Gram to the
// application inside mcm_transform_moments, so that the 8 GB particle matrix is not read a third time for them.
struct StatsHandoff {
  bool armed = false, valid = false;
  int M = 0;
  size_t n = 0;
  const void* first_var = nullptr;
  int* d_ex = nullptr;
  double* d_ratio = nullptr;
};
StatsHandoff& stats_handoff();
}  // namespace mcm
#endif  // !__CUDACC_RTC__

#ifdef __CUDACC__
namespace mcm {

// 128-bit pack of T (2 x f64 or 4 x f32) — the unit of every global access.
template <typename T>
struct alignas(16) Pack {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

// Streaming (evict-first) 128-bit loads/stores: particle arrays at N=1e8 are
// read once and written once per launch; nothing is worth keeping in L1/L2.
template <typename T>
__device__ __forceinline__ Pack<T> ld_stream(const Pack<T>* p) {
  Pack<T> r;
  const int4 raw = __ldcs(reinterpret_cast<const int4*>(p));
  *reinterpret_cast<int4*>(&r) = raw;
  return r;
}
template <typename T>
__device__ __forceinline__ void st_stream(Pack<T>* p, const Pack<T>& v) {
  __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&v));
}

// NaN test on the integer pipe (the FP64 pipe is the scarce resource of Float64 chains): exponent all
// ones and a non-zero HIGH mantissa word — exact for every NaN arithmetic can produce (quiet NaNs).
__device__ __forceinline__ bool is_nan_fast(double v) { return ((unsigned)__double2hiint(v) << 1) > 0xffe00000u; }
__device__ __forceinline__ bool is_nan_fast(float v) { return (__float_as_uint(v) << 1) > 0xff000000u; }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

}  // namespace mcm
#endif
