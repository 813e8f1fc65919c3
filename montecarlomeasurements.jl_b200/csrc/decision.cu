// decision.cu — the remaining device pieces of the decision layer:
//   with_nominal (src/nominal.jl:1-29): argmin_i |x_i - val| (scalar) / argmin_i sum_j (x_ij - val_j)^2 (joint),
//                                       then shuffle_and_insert: out = copy(in); out[ind] = in[0]; out[0] = val;
//   intersect    (src/particles.jl:375-380): stable filter lo <= x <= hi of both operands, concatenated.
// (union, particles.jl:367-369, is two device-to-device copies and lives on the host side.)
// argmin follows Julia's findmin: the FIRST index of the minimum, and a NaN distance beats every number.
#include <math.h>
#include <string.h>
#include "mcm_common.cuh"

namespace mcm {

struct Best {
  double d;                 // distance; NaN allowed
  unsigned long long i;     // index
};
__device__ __forceinline__ bool best_less(const Best& a, const Best& b) {   // a strictly preferable to b
  const bool an = a.d != a.d, bn = b.d != b.d;
  if (an != bn) return an;
  if (!an && a.d != b.d) return a.d < b.d;
  return a.i < b.i;
}
__device__ __forceinline__ Best best_block_reduce(Best v) {
  __shared__ Best s_w[kThreads / 32];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    Best o;
    o.d = __shfl_down_sync(0xffffffffu, v.d, d);
    o.i = __shfl_down_sync(0xffffffffu, v.i, d);
    if (best_less(o, v)) v = o;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kThreads / 32 ? s_w[lane] : Best{__builtin_huge_val(), ~0ull};
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) {
      Best o;
      o.d = __shfl_down_sync(0xffffffffu, v.d, d);
      o.i = __shfl_down_sync(0xffffffffu, v.i, d);
      if (best_less(o, v)) v = o;
    }
  }
  __syncthreads();
  return v;
}

constexpr int AM_MAXVARS = 64;     // variables per pointer table slab of the joint form
struct VarTable { const void* p[AM_MAXVARS]; double val[AM_MAXVARS]; };

// M == 0: scalar form on `x`; else joint form accumulating into dist[] slab by slab (first: overwrite)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_absdiff_best(const T* __restrict__ x, size_t n, double val, Best* __restrict__ part) {
  Best b{__builtin_huge_val(), ~0ull};
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const Best c{fabs((double)x[i] - val), (unsigned long long)i};
    if (best_less(c, b)) b = c;
  }
  b = best_block_reduce(b);
  if (threadIdx.x == 0) part[blockIdx.x] = b;
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_sqdist_accum(const __grid_constant__ VarTable tab, int m, size_t n, double* __restrict__ dist, int first) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    double acc = first ? 0.0 : dist[i];
    for (int j = 0; j < m; ++j) {                       // Julia's sum(abs2, ...) over a row: left to right
      const double d = (double)static_cast<const T*>(tab.p[j])[i] - tab.val[j];
      acc += d * d;
    }
    dist[i] = acc;
  }
}
__global__ void __launch_bounds__(kThreads)
k_best_of(const double* __restrict__ dist, size_t n, Best* __restrict__ part) {
  Best b{__builtin_huge_val(), ~0ull};
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const Best c{dist[i], (unsigned long long)i};
    if (best_less(c, b)) b = c;
  }
  b = best_block_reduce(b);
  if (threadIdx.x == 0) part[blockIdx.x] = b;
}
__global__ void __launch_bounds__(kThreads) k_best_final(const Best* __restrict__ part, int nblk, Best* __restrict__ out) {
  Best b{__builtin_huge_val(), ~0ull};
  for (int i = threadIdx.x; i < nblk; i += kThreads)
    if (best_less(part[i], b)) b = part[i];
  b = best_block_reduce(b);
  if (threadIdx.x == 0) *out = b;
}

template <typename T>
__global__ void k_shuffle_insert(const T* __restrict__ in, T* __restrict__ out, unsigned long long ind, double val) {
  const T first = in[0];              // read before either write: in may alias out
  out[ind] = first;
  out[0] = (T)val;
}

// ---- stable range filter ---------------------------------------------------------------------------
constexpr int FL_IPT = 8;                       // consecutive elements per thread (blocked => order-preserving)
constexpr int FL_TILE = kThreads * FL_IPT;

template <typename T>
__device__ __forceinline__ unsigned filter_mask(const T* __restrict__ in, size_t n, size_t base, double lo, double hi, T (&v)[FL_IPT]) {
  unsigned mask = 0;
#pragma unroll
  for (int k = 0; k < FL_IPT; ++k) {
    const size_t i = base + k;
    if (i < n) {
      v[k] = in[i];
      const double x = (double)v[k];
      if (lo <= x && x <= hi) mask |= 1u << k;    // false for NaN, as in the reference's filter
    }
  }
  return mask;
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_filter_count(const T* __restrict__ in, size_t n, double lo, double hi, unsigned long long* __restrict__ counts) {
  T v[FL_IPT];
  const size_t base = (size_t)blockIdx.x * FL_TILE + (size_t)threadIdx.x * FL_IPT;
  const int c = __popc(filter_mask(in, n, base, lo, hi, v));
  __shared__ int s_w[kThreads / 32];
  int w = c;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) w += __shfl_down_sync(0xffffffffu, w, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < kThreads / 32; ++k) t += s_w[k];
    counts[blockIdx.x] = (unsigned long long)t;
  }
}
// exclusive scan of the per-tile counts in place, by one CTA; counts[ntiles] receives the total
__global__ void __launch_bounds__(1024) k_filter_scan(unsigned long long* __restrict__ counts, size_t ntiles) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t b = 0; b < ntiles; b += 1024) {
    const size_t i = b + threadIdx.x;
    const unsigned long long v = i < ntiles ? counts[i] : 0ull;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned long long t = s_w[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t += o;
      }
      s_w[lane] = t;          // inclusive over warps
    }
    __syncthreads();
    const unsigned long long carry = s_carry;
    const unsigned long long before = carry + (warp ? s_w[warp - 1] : 0ull) + (inc - v);
    if (i < ntiles) counts[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[ntiles] = s_carry;
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_filter_write(const T* __restrict__ in, size_t n, double lo, double hi, const unsigned long long* __restrict__ offs,
               T* __restrict__ out) {
  T v[FL_IPT];
  const size_t base = (size_t)blockIdx.x * FL_TILE + (size_t)threadIdx.x * FL_IPT;
  const unsigned mask = filter_mask(in, n, base, lo, hi, v);
  const int c = __popc(mask);
  __shared__ int s_w[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  int wbase = 0;
  for (int k = 0; k < warp; ++k) wbase += s_w[k];
  size_t pos = (size_t)offs[blockIdx.x] + wbase + (inc - c);
#pragma unroll
  for (int k = 0; k < FL_IPT; ++k)
    if (mask & (1u << k)) out[pos++] = v[k];
}

int read_back(double* host_dst, const double* dev_src, size_t count);

template <typename T>
static int best_scalar(const T* x, size_t n, double val, Best* d_part, int grid) {
  k_absdiff_best<T><<<grid, kThreads, 0, ctx().stream>>>(x, n, val, d_part);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

static int best_finish(Best* d_part, int grid, uint64_t* idx_host) {
  Ctx& c = ctx();
  k_best_final<<<1, kThreads, 0, c.stream>>>(d_part, grid, d_part + grid);
  MCM_CHECK_LAUNCH();
  double h[2];
  int rc = read_back(h, reinterpret_cast<const double*>(d_part + grid), 2);
  if (rc) return rc;
  unsigned long long i;
  memcpy(&i, &h[1], 8);
  *idx_host = (uint64_t)i;
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

static int am_grid(size_t n) {
  const size_t want = (n + kThreads - 1) / kThreads;
  const size_t cap = (size_t)ctx().sm_count * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

extern "C" int mcm_argmin_absdiff(int dtype, const void* in, size_t n, double val, uint64_t* idx_host) {
  MCM_NEED_INIT();
  if (!in || !idx_host || n == 0) return fail(MCM_ERR_INVALID, "mcm_argmin_absdiff: bad arguments");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_argmin_absdiff: bad dtype %d", dtype);
  Ctx& c = ctx();
  const int grid = am_grid(n);
  Best* d_part = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_part, (size_t)(grid + 1) * sizeof(Best), c.stream));
  int rc = dtype == MCM_F64 ? best_scalar<double>((const double*)in, n, val, d_part, grid)
                            : best_scalar<float>((const float*)in, n, val, d_part, grid);
  if (!rc) rc = best_finish(d_part, grid, idx_host);
  cudaFreeAsync(d_part, c.stream);
  return rc;
}

extern "C" int mcm_argmin_sqdist(int dtype, const void* const* vars, int M, size_t n, const double* val_host,
                                 uint64_t* idx_host) {
  MCM_NEED_INIT();
  if (!vars || !val_host || !idx_host || M < 1 || n == 0) return fail(MCM_ERR_INVALID, "mcm_argmin_sqdist: bad arguments");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_argmin_sqdist: bad dtype %d", dtype);
  for (int j = 0; j < M; ++j)
    if (!vars[j]) return fail(MCM_ERR_INVALID, "mcm_argmin_sqdist: variable %d is null", j);
  Ctx& c = ctx();
  const int grid = am_grid(n);
  double* d_dist = nullptr;
  Best* d_part = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_dist, n * sizeof(double), c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_part, (size_t)(grid + 1) * sizeof(Best), c.stream));
  int rc = MCM_OK;
  for (int j0 = 0; j0 < M && !rc; j0 += AM_MAXVARS) {
    VarTable tab;
    const int m = M - j0 < AM_MAXVARS ? M - j0 : AM_MAXVARS;
    for (int j = 0; j < m; ++j) { tab.p[j] = vars[j0 + j]; tab.val[j] = val_host[j0 + j]; }
    if (dtype == MCM_F64) k_sqdist_accum<double><<<grid, kThreads, 0, c.stream>>>(tab, m, n, d_dist, j0 == 0);
    else k_sqdist_accum<float><<<grid, kThreads, 0, c.stream>>>(tab, m, n, d_dist, j0 == 0);
    if (cudaGetLastError() != cudaSuccess) rc = fail(MCM_ERR_CUDA, "mcm_argmin_sqdist: launch failed");
  }
  if (!rc) {
    k_best_of<<<grid, kThreads, 0, c.stream>>>(d_dist, n, d_part);
    if (cudaGetLastError() != cudaSuccess) rc = fail(MCM_ERR_CUDA, "mcm_argmin_sqdist: launch failed");
  }
  if (!rc) rc = best_finish(d_part, grid, idx_host);
  cudaFreeAsync(d_dist, c.stream);
  cudaFreeAsync(d_part, c.stream);
  return rc;
}

extern "C" int mcm_shuffle_insert(int dtype, const void* in, size_t n, uint64_t ind, double val, void* out) {
  MCM_NEED_INIT();
  if (!in || !out || n == 0 || ind >= n) return fail(MCM_ERR_INVALID, "mcm_shuffle_insert: bad arguments");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_shuffle_insert: bad dtype %d", dtype);
  Ctx& c = ctx();
  const size_t es = dtype == MCM_F64 ? 8 : 4;
  if (in != out) MCM_CUDA(cudaMemcpyAsync(out, in, n * es, cudaMemcpyDeviceToDevice, c.stream));
  if (dtype == MCM_F64) k_shuffle_insert<double><<<1, 1, 0, c.stream>>>((const double*)in, (double*)out, ind, val);
  else k_shuffle_insert<float><<<1, 1, 0, c.stream>>>((const float*)in, (float*)out, ind, val);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

extern "C" int mcm_filter_range(int dtype, const void* in, size_t n, double lo, double hi, void* out, size_t out_offset,
                                size_t* count_host) {
  MCM_NEED_INIT();
  if (!count_host || (n && (!in || !out))) return fail(MCM_ERR_INVALID, "mcm_filter_range: bad arguments");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_filter_range: bad dtype %d", dtype);
  *count_host = 0;
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  const size_t ntiles = (n + FL_TILE - 1) / FL_TILE;
  if (ntiles > 0x7fffffffull) return fail(MCM_ERR_UNSUPPORTED, "mcm_filter_range: n too large");
  unsigned long long* d_counts = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_counts, (ntiles + 1) * sizeof(unsigned long long), c.stream));
  int rc = MCM_OK;
  if (dtype == MCM_F64) k_filter_count<double><<<(unsigned)ntiles, kThreads, 0, c.stream>>>((const double*)in, n, lo, hi, d_counts);
  else k_filter_count<float><<<(unsigned)ntiles, kThreads, 0, c.stream>>>((const float*)in, n, lo, hi, d_counts);
  k_filter_scan<<<1, 1024, 0, c.stream>>>(d_counts, ntiles);
  if (dtype == MCM_F64)
    k_filter_write<double><<<(unsigned)ntiles, kThreads, 0, c.stream>>>((const double*)in, n, lo, hi, d_counts, (double*)out + out_offset);
  else
    k_filter_write<float><<<(unsigned)ntiles, kThreads, 0, c.stream>>>((const float*)in, n, lo, hi, d_counts, (float*)out + out_offset);
  if (cudaGetLastError() != cudaSuccess) rc = fail(MCM_ERR_CUDA, "mcm_filter_range: launch failed");
  if (!rc) {
    double h;
    rc = read_back(&h, reinterpret_cast<const double*>(d_counts + ntiles), 1);
    unsigned long long t;
    memcpy(&t, &h, 8);
    *count_host = (size_t)t;
  }
  cudaFreeAsync(d_counts, c.stream);
  return rc;
}
