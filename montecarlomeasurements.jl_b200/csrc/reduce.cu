// reduce.cu — pmean / pvar / pstd / pminimum / pmaximum / sum / @prob counts.
// Reference: src/particles.jl:302-313 (pvar, pstd; N==1 -> zero(T)),
// :316-323 (pmean, pextrema, pminimum, pmaximum forwarded to Statistics/Base),
// src/bymap.jl:128-136 (@prob == mean(p.particles .< a)).
#include <math.h>
#include <string.h>
#include "mcm_moments.cuh"
#include "mcm_bulk.cuh"
#include "mcm_math.cuh"

namespace mcm {

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_moments(const T* __restrict__ in, size_t n, Mom* partials, unsigned int* ticket, double* out,
          const __grid_constant__ MomExchange px) {
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int UNROLL = 4;
  constexpr int C = V * UNROLL;
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const size_t nvec = vec ? n / V : 0;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const size_t rows_full = nvec / (stride * UNROLL);
  const P* pin = reinterpret_cast<const P*>(in);

  Mom m = mom_empty();
  for (size_t r = 0; r < rows_full; ++r) {
    const size_t base = r * stride * UNROLL + tid;
    P x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) x[u] = ld_stream(pin + base + (size_t)u * stride);
    double v[C];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int e = 0; e < V; ++e) v[u * V + e] = (double)x[u].v[e];
    mom_push_chunk<C>(m, v);
  }
  for (size_t j = rows_full * stride * UNROLL + tid; j < nvec; j += stride) {
    const P x = ld_stream(pin + j);
    double v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = (double)x.v[e];
    mom_push_chunk<V>(m, v);
  }
  for (size_t i = nvec * V + tid; i < n; i += stride) mom_push_one(m, (double)in[i]);

  m = mom_block_reduce(m);
  mom_grid_finish(m, partials, ticket, out, &px);
}

// ---- TMA (bulk async copy) variant for large inputs ---------------------------------------
// One elected thread streams 16 KiB tiles global -> shared with cp.async.bulk, completion tracked by
// an mbarrier per stage (SASS: UBLKCP + SYNCS); all 256 threads then read the tile from shared
// memory with conflict-free 128-bit LDS.  Bytes in flight per SM no longer depend on registers or
// occupancy: 3 resident CTAs x 4 stages x 16 KiB.  Same per-thread chunk statistics as k_moments.
constexpr int MB_STAGES = 4;
constexpr int MB_TILE_BYTES = 16384;

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_moments_bulk(const T* __restrict__ in, size_t n, Mom* partials, unsigned int* ticket, double* out,
               const __grid_constant__ MomExchange px) {
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int EPT = MB_TILE_BYTES / (int)sizeof(T);           // elements per tile
  constexpr int PPT = MB_TILE_BYTES / 16 / kThreads;            // packs per thread per tile (4)
  constexpr int C = PPT * V;                                    // chunk size (8 doubles / 16 floats)
  extern __shared__ __align__(128) unsigned char smem_tiles[];
  __shared__ __align__(8) unsigned long long full_bar[MB_STAGES];
  const int tid = threadIdx.x;
  const size_t ntiles = n / EPT;
  const size_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < MB_STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < MB_STAGES && (size_t)s < my_tiles; ++s) {
      const size_t tile = blockIdx.x + (size_t)s * gridDim.x;
      mbar_expect_tx(&full_bar[s], MB_TILE_BYTES);
      bulk_g2s(smem_tiles + (size_t)s * MB_TILE_BYTES, in + tile * EPT, MB_TILE_BYTES, &full_bar[s]);
    }
  }
  Mom m = mom_empty();
  for (size_t i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % MB_STAGES);
    const unsigned parity = (unsigned)((i / MB_STAGES) & 1);
    mbar_wait(&full_bar[s], parity);
    const P* tile = reinterpret_cast<const P*>(smem_tiles + (size_t)s * MB_TILE_BYTES);
    double v[C];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const P x = tile[k * kThreads + tid];
#pragma unroll
      for (int e = 0; e < V; ++e) v[k * V + e] = (double)x.v[e];
    }
    mom_push_chunk<C>(m, v);
    __syncthreads();  // every thread has read stage s: it may be refilled
    if (tid == 0 && i + MB_STAGES < my_tiles) {
      const size_t tile_next = blockIdx.x + (i + MB_STAGES) * gridDim.x;
      mbar_expect_tx(&full_bar[s], MB_TILE_BYTES);
      bulk_g2s(smem_tiles + (size_t)s * MB_TILE_BYTES, in + tile_next * EPT, MB_TILE_BYTES, &full_bar[s]);
    }
  }
  // remainder (< one tile) through the generic path
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = ntiles * EPT + (size_t)blockIdx.x * kThreads + tid; j < n; j += stride) mom_push_one(m, (double)in[j]);
  m = mom_block_reduce(m);
  mom_grid_finish(m, partials, ticket, out, &px);
}

template <typename T, int CMP>
__global__ void __launch_bounds__(kThreads)
k_count(const T* __restrict__ in, size_t n, T thr, unsigned long long* out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  unsigned long long c = 0;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const T x = in[i];
    bool b;
    if constexpr (CMP == MCM_CMP_LT) b = x < thr;
    else if constexpr (CMP == MCM_CMP_LE) b = x <= thr;
    else if constexpr (CMP == MCM_CMP_GT) b = x > thr;
    else if constexpr (CMP == MCM_CMP_GE) b = x >= thr;
    else if constexpr (CMP == MCM_CMP_EQ) b = x == thr;
    else if constexpr (CMP == MCM_CMP_NE) b = x != thr;
    else if constexpr (CMP == MCM_CMP_ISINTEGER) b = (x - x == T(0)) && (x == (sizeof(T) == 8 ? (T)trunc((double)x) : (T)truncf((float)x)));
    else if constexpr (CMP == MCM_CMP_ISFINITE) b = (x - x == T(0));
    else b = x != x;
    c += b ? 1ull : 0ull;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// a[i] cmp b[i]: p1 == p2 compares the particle vectors (particles.jl:488); :montecarlo comparisons (:512-515)
template <typename T, int CMP>
__global__ void __launch_bounds__(kThreads)
k_count2(const T* __restrict__ a, const T* __restrict__ b, size_t n, unsigned long long* out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  unsigned long long c = 0;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const T x = a[i], y = b[i];
    bool r;
    if constexpr (CMP == MCM_CMP_LT) r = x < y;
    else if constexpr (CMP == MCM_CMP_LE) r = x <= y;
    else if constexpr (CMP == MCM_CMP_GT) r = x > y;
    else if constexpr (CMP == MCM_CMP_GE) r = x >= y;
    else if constexpr (CMP == MCM_CMP_EQ) r = x == y;
    else r = x != y;
    c += r ? 1ull : 0ull;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// round(Int, p, r) (particles.jl:347-348): Int64(round(x, r)); InexactError -> DOMAIN flag, INT64_MIN stored
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_round_int(const T* __restrict__ in, int mode, long long* __restrict__ out, size_t n, uint32_t* flags) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const double x = (double)in[i];
    double r;
    switch (mode) {
      case MCM_RM_NEAREST: r = rint(x); break;
      case MCM_RM_TOZERO: r = trunc(x); break;
      case MCM_RM_UP: r = ceil(x); break;
      case MCM_RM_DOWN: r = floor(x); break;
      case MCM_RM_TIESAWAY: r = math::m_round_tiesaway(in[i]); break;
      case MCM_RM_TIESUP: r = math::m_round_tiesup(in[i]); break;
      default: r = math::m_round_fromzero(in[i]); break;
    }
    if (r >= -9223372036854775808.0 && r < 9223372036854775808.0) out[i] = (long long)r;
    else { out[i] = (long long)0x8000000000000000ull; bad = true; }
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
}

static int g_use_bulk = 1;  // mcm_moments_use_bulk(0) forces the LDG kernel (A/B measurements, tests)

MomExchange mom_exchange_none() {
  MomExchange px;
  memset(&px, 0, sizeof(px));
  px.world = 1;
  px.flags = ctx().d_flags;
  return px;
}

// px == nullptr: this rank only; else the partial is merged with every other rank's inside the kernel (comm.cu)
int moments_device(int dtype, const void* in, size_t n, double* d_out6, const MomExchange* pxp) {
  Ctx& c = ctx();
  const MomExchange px = pxp ? *pxp : mom_exchange_none();
  static int occ64 = 0, occ32 = 0;
  if (occ64 == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, k_moments<double>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, k_moments<float>, kThreads, 0);
    if (occ64 < 1) occ64 = 1;
    if (occ32 < 1) occ32 = 1;
  }
  Mom* partials = reinterpret_cast<Mom*>(c.d_scratch);
  const size_t max_grid = c.scratch_doubles / 8;  // 7 doubles per partial, keep headroom
  // large aligned inputs: bulk-async (TMA) pipeline
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  if (g_use_bulk && (dtype == MCM_F64 || dtype == MCM_F32) && n * esz >= ((size_t)8 << 20) &&
      (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
    constexpr size_t kSmem = (size_t)MB_STAGES * MB_TILE_BYTES;
    static int occb = 0;
    if (occb == 0) {
      MCM_CUDA(cudaFuncSetAttribute(k_moments_bulk<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
      MCM_CUDA(cudaFuncSetAttribute(k_moments_bulk<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occb, k_moments_bulk<double>, kThreads, kSmem);
      if (occb < 1) occb = 1;
      if (occb > 3) occb = 3;
    }
    int grid = c.sm_count * occb;
    const size_t ntiles = n * esz / MB_TILE_BYTES;
    if ((size_t)grid > ntiles) grid = (int)(ntiles ? ntiles : 1);
    if (dtype == MCM_F64)
      k_moments_bulk<double><<<grid, kThreads, kSmem, c.stream>>>((const double*)in, n, partials, c.d_tickets, d_out6, px);
    else
      k_moments_bulk<float><<<grid, kThreads, kSmem, c.stream>>>((const float*)in, n, partials, c.d_tickets, d_out6, px);
    MCM_CHECK_LAUNCH();
    return MCM_OK;
  }
  if (dtype == MCM_F64) {
    int grid = stream_grid((n + 1) / 2, 4, occ64 > 4 ? 4 : occ64);
    if ((size_t)grid > max_grid) grid = (int)max_grid;
    k_moments<double><<<grid, kThreads, 0, c.stream>>>((const double*)in, n, partials, c.d_tickets, d_out6, px);
  } else if (dtype == MCM_F32) {
    int grid = stream_grid((n + 3) / 4, 4, occ32 > 4 ? 4 : occ32);
    if ((size_t)grid > max_grid) grid = (int)max_grid;
    k_moments<float><<<grid, kThreads, 0, c.stream>>>((const float*)in, n, partials, c.d_tickets, d_out6, px);
  } else {
    return fail(MCM_ERR_INVALID, "moments: bad dtype %d", dtype);
  }
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

int moments_device(int dtype, const void* in, size_t n, double* d_out6) { return moments_device(dtype, in, n, d_out6, nullptr); }

// device result slot (6 doubles) lives at the end of the scratch buffer
double* moments_result_slot() { return ctx().d_scratch + (ctx().scratch_doubles - 16); }

// ---- result ring -----------------------------------------------------------------------------------------------
// Statistics that go back to the caller are written by the kernel's last CTA straight into MAPPED PINNED host
// memory (no cudaMemcpy), eight slots used round-robin, one event per slot: an asynchronous launch returns a
// ticket, and waiting for ticket k (cudaEventSynchronize) does not wait for launches queued after it — step k's
// readback overlaps kernel k+1.
struct ResultRing {
  double* h = nullptr;              // 8 x 8 doubles, cudaHostAllocMapped (same pointer on the device: UVA)
  cudaEvent_t ev[8];
  unsigned long long seq = 0;       // tickets handed out so far
};
static ResultRing g_ring;
int ring_acquire(double** dev_slot, unsigned long long* ticket) {
  ResultRing& r = g_ring;
  if (!r.h) {
    MCM_CUDA(cudaHostAlloc((void**)&r.h, 8 * 8 * sizeof(double), cudaHostAllocMapped | cudaHostAllocPortable));
    for (int i = 0; i < 8; ++i) MCM_CUDA(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
  }
  *ticket = r.seq++;
  *dev_slot = r.h + 8 * (*ticket % 8);
  return MCM_OK;
}
int ring_publish(unsigned long long ticket) {
  MCM_CUDA(cudaEventRecord(g_ring.ev[ticket % 8], ctx().stream));
  return MCM_OK;
}
int ring_wait(unsigned long long ticket, double* out6) {
  ResultRing& r = g_ring;
  if (!r.h || ticket >= r.seq || r.seq - ticket > 8)
    return fail(MCM_ERR_INVALID, "statistics ticket %llu is not among the last 8 launches", ticket);
  MCM_CUDA(cudaEventSynchronize(r.ev[ticket % 8]));
  const volatile double* src = r.h + 8 * (ticket % 8);
  for (int i = 0; i < 6; ++i) out6[i] = src[i];
  return MCM_OK;
}
unsigned long long ring_last_ticket() { return g_ring.seq ? g_ring.seq - 1 : 0; }
double* ring_host_slot(unsigned long long ticket) { return g_ring.h + 8 * (ticket % 8); }

int read_back(double* host_dst, const double* dev_src, size_t count) {
  Ctx& c = ctx();
  if (count * sizeof(double) > c.pinned_bytes) return fail(MCM_ERR_INVALID, "read_back: too large");
  MCM_CUDA(cudaMemcpyAsync(c.h_pinned, dev_src, count * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  for (size_t i = 0; i < count; ++i) host_dst[i] = static_cast<double*>(c.h_pinned)[i];
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_moments_use_bulk(int on) { g_use_bulk = on; return MCM_OK; }

extern "C" int mcm_moments(int dtype, const void* in, size_t n, double* out_host) {
  MCM_NEED_INIT();
  if (!out_host || (n && !in)) return fail(MCM_ERR_INVALID, "mcm_moments: null pointer");
  if (n == 0) {
    Mom e = mom_empty();
    out_host[0] = 0; out_host[1] = NAN; out_host[2] = 0; out_host[3] = e.mn; out_host[4] = e.mx; out_host[5] = 0;
    return MCM_OK;
  }
  double* slot = moments_result_slot();
  int rc = moments_device(dtype, in, n, slot);
  if (rc) return rc;
  return read_back(out_host, slot, 6);
}

// Asynchronous form: the kernel is enqueued and its six moments stay in a device slot until mcm_moments_collect reads the
// slot of the LAST enqueue (one read-back for a burst of reductions whose results are consumed later, and the way to time
// the kernel without the host round trip every synchronous call pays).
extern "C" int mcm_moments_enqueue(int dtype, const void* in, size_t n) {
  MCM_NEED_INIT();
  if (!in || n == 0) return fail(MCM_ERR_INVALID, "mcm_moments_enqueue: null pointer or no particles");
  return moments_device(dtype, in, n, moments_result_slot());
}
extern "C" int mcm_moments_collect(double* out_host) {
  MCM_NEED_INIT();
  if (!out_host) return fail(MCM_ERR_INVALID, "mcm_moments_collect: null pointer");
  return read_back(out_host, moments_result_slot(), 6);
}

extern "C" int mcm_moments_merge(const double* a, const double* b, double* out) {
  if (!a || !b || !out) return fail(MCM_ERR_INVALID, "mcm_moments_merge: null pointer");
  Mom ma{a[0], a[1], 0.0, a[2], a[3], a[4], a[5]}, mb{b[0], b[1], 0.0, b[2], b[3], b[4], b[5]};
  Mom r = mom_merge(ma, mb);
  out[0] = r.n; out[1] = r.K + r.dm; out[2] = r.m2; out[3] = r.mn; out[4] = r.mx; out[5] = r.sum;
  return MCM_OK;
}

// kind -> scalar from the six moments; shared with the sharded path
extern "C" int mcm_reduce_from_moments(int kind, const double* m, double* out) {
  const double n = m[0];
  switch (kind) {
    case MCM_R_SUM: *out = m[5]; break;
    case MCM_R_MEAN: *out = m[1]; break;
    // particles.jl:308-312: N == 1 -> zero(T)
    case MCM_R_VAR: *out = (n == 1.0) ? 0.0 : m[2] / (n - 1.0); break;
    case MCM_R_VARP: *out = (n == 1.0) ? 0.0 : m[2] / n; break;
    case MCM_R_STD: *out = (n == 1.0) ? 0.0 : sqrt(m[2] / (n - 1.0)); break;
    case MCM_R_STDP: *out = (n == 1.0) ? 0.0 : sqrt(m[2] / n); break;
    case MCM_R_MIN: *out = m[3]; break;
    case MCM_R_MAX: *out = m[4]; break;
    default: return fail(MCM_ERR_INVALID, "mcm_reduce: unknown kind %d", kind);
  }
  return MCM_OK;
}

extern "C" int mcm_reduce(int kind, int dtype, const void* in, size_t n, double* out_host) {
  MCM_NEED_INIT();
  if (!out_host) return fail(MCM_ERR_INVALID, "mcm_reduce: null output");
  if (n == 0) return fail(MCM_ERR_INVALID, "mcm_reduce: empty particles (ArgumentError in Base for mean/extrema of empty)");
  double m[6];
  int rc = mcm_moments(dtype, in, n, m);
  if (rc) return rc;
  return mcm_reduce_from_moments(kind, m, out_host);
}

template <typename T>
static int run_count(int cmp, const T* in, size_t n, T thr, unsigned long long* d_out) {
  Ctx& c = ctx();
  const int grid = stream_grid(n, 8, 8);
  switch (cmp) {
#define MCM_CASE(K) \
  case K: k_count<T, K><<<grid, kThreads, 0, c.stream>>>(in, n, thr, d_out); break;
    MCM_CASE(MCM_CMP_LT) MCM_CASE(MCM_CMP_LE) MCM_CASE(MCM_CMP_GT) MCM_CASE(MCM_CMP_GE)
    MCM_CASE(MCM_CMP_EQ) MCM_CASE(MCM_CMP_NE) MCM_CASE(MCM_CMP_ISINTEGER) MCM_CASE(MCM_CMP_ISFINITE)
    MCM_CASE(MCM_CMP_ISNAN)
#undef MCM_CASE
    default: return fail(MCM_ERR_INVALID, "mcm_count: unknown comparison %d", cmp);
  }
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

template <typename T>
static int run_count2(int cmp, const T* a, const T* b, size_t n, unsigned long long* d_out) {
  Ctx& c = ctx();
  const int grid = stream_grid(n, 8, 8);
  switch (cmp) {
#define MCM_CASE(K) \
  case K: k_count2<T, K><<<grid, kThreads, 0, c.stream>>>(a, b, n, d_out); break;
    MCM_CASE(MCM_CMP_LT) MCM_CASE(MCM_CMP_LE) MCM_CASE(MCM_CMP_GT) MCM_CASE(MCM_CMP_GE)
    MCM_CASE(MCM_CMP_EQ) MCM_CASE(MCM_CMP_NE)
#undef MCM_CASE
    default: return fail(MCM_ERR_INVALID, "mcm_count2: unknown comparison %d", cmp);
  }
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

extern "C" int mcm_count(int cmp, int dtype, const void* in, size_t n, double threshold, uint64_t* count_host) {
  MCM_NEED_INIT();
  if (!count_host || (n && !in)) return fail(MCM_ERR_INVALID, "mcm_count: null pointer");
  Ctx& c = ctx();
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(moments_result_slot() + 8);
  MCM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), c.stream));
  int rc = MCM_OK;
  if (n) {
    if (dtype == MCM_F64) rc = run_count<double>(cmp, (const double*)in, n, threshold, d_out);
    else if (dtype == MCM_F32) rc = run_count<float>(cmp, (const float*)in, n, (float)threshold, d_out);
    else return fail(MCM_ERR_INVALID, "mcm_count: bad dtype %d", dtype);
  }
  if (rc) return rc;
  double tmp;
  rc = read_back(&tmp, reinterpret_cast<const double*>(d_out), 1);
  if (rc) return rc;
  uint64_t bits;
  memcpy(&bits, &tmp, 8);
  *count_host = bits;
  return MCM_OK;
}

extern "C" int mcm_count2(int cmp, int dtype, const void* a, const void* b, size_t n, uint64_t* count_host) {
  MCM_NEED_INIT();
  if (!count_host || (n && (!a || !b))) return fail(MCM_ERR_INVALID, "mcm_count2: null pointer");
  Ctx& c = ctx();
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(moments_result_slot() + 8);
  MCM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), c.stream));
  int rc = MCM_OK;
  if (n) {
    if (dtype == MCM_F64) rc = run_count2<double>(cmp, (const double*)a, (const double*)b, n, d_out);
    else if (dtype == MCM_F32) rc = run_count2<float>(cmp, (const float*)a, (const float*)b, n, d_out);
    else return fail(MCM_ERR_INVALID, "mcm_count2: bad dtype %d", dtype);
  }
  if (rc) return rc;
  double tmp;
  rc = read_back(&tmp, reinterpret_cast<const double*>(d_out), 1);
  if (rc) return rc;
  uint64_t bits;
  memcpy(&bits, &tmp, 8);
  *count_host = bits;
  return MCM_OK;
}

extern "C" int mcm_round_int(int dtype, const void* in, int mode, int64_t* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_round_int: null pointer");
  if (mode < MCM_RM_NEAREST || mode > MCM_RM_FROMZERO) return fail(MCM_ERR_INVALID, "mcm_round_int: unknown rounding mode %d", mode);
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  const int grid = stream_grid(n, 4, 8);
  if (dtype == MCM_F64) k_round_int<double><<<grid, kThreads, 0, c.stream>>>((const double*)in, mode, (long long*)out, n, c.d_flags);
  else if (dtype == MCM_F32) k_round_int<float><<<grid, kThreads, 0, c.stream>>>((const float*)in, mode, (long long*)out, n, c.d_flags);
  else return fail(MCM_ERR_INVALID, "mcm_round_int: bad dtype %d", dtype);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
