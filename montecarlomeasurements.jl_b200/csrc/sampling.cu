// sampling.cu — systematic_sample (src/sampling.jl:6-17), bootstrap
// (src/resampling.jl:76-84) and systematic resampling (the dead reference code at
// src/resampling.jl:45-67, restated with an exact integer scan).
#include <math.h>
#include "mcm_common.cuh"
#include "mcm_norminv.cuh"

namespace mcm {

// ---- systematic sample --------------------------------------------------------
// sampling.jl:8-10: e = 0.5/N; y = e:1/N:1; o = quantile.(d, y).  Grid point i
// (0-based) is y_i = (2i+1)/(2N), evaluated as one correctly rounded division
// (see DESIGN.md H3 for how this relates to Julia's StepRangeLen).
// quantile(Normal(mu,sigma), p) = mu + sigma * norminvcdf(p)  [StatsFuns].
template <typename T, bool NORMAL>
__global__ void __launch_bounds__(kThreads)
k_systematic(T* __restrict__ out, size_t n, double a, double b, const int64_t* __restrict__ perm,
             unsigned long long i0, unsigned long long n_total) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  const double two_n = 2.0 * (double)n_total;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    const unsigned long long i = perm ? (unsigned long long)perm[j] : i0 + j;
    const double p = (double)(2ull * i + 1ull) / two_n;
    double v;
    if constexpr (NORMAL) v = a + b * norminv::norminvcdf(p);
    else v = a + (b - a) * p;  // quantile(Uniform(a,b), p)
    out[j] = (T)v;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_norminv(const T* __restrict__ p, T* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride)
    out[j] = (T)norminv::norminvcdf((double)p[j]);
}

// ---- Philox-4x32-10 (counter-based; Salmon et al. 2011) -------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform index in [0,N): 64 random bits, multiply-high (Lemire, without rejection:
// bias < N/2^64 <= 2^-37 for N <= 2^27, far below Monte-Carlo resolution)
__global__ void __launch_bounds__(kThreads)
k_rand_indices(int64_t* __restrict__ out, size_t n, unsigned long long N, unsigned long long seed,
               unsigned long long offset) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    uint32_t r[4];
    philox4x32(offset + j, seed, r);
    const unsigned long long u = ((unsigned long long)r[1] << 32) | r[0];
    out[j] = (int64_t)__umul64hi(u, N);
  }
}
__global__ void __launch_bounds__(kThreads)
k_rand_keys(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, size_t n, unsigned long long seed) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    uint32_t r[4];
    philox4x32(j, seed, r);
    keys[j] = ((unsigned long long)r[1] << 32) | r[0];
    vals[j] = (uint32_t)j;
  }
}
__global__ void __launch_bounds__(256) k_u32_to_i64_b(const uint32_t* __restrict__ v, int64_t* __restrict__ o, size_t n) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) o[i] = (int64_t)v[i];
}

// ---- gather ----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_gather(const T* __restrict__ src, const int64_t* __restrict__ idx, T* __restrict__ out, size_t n_draw,
         size_t n_src, uint32_t* flags) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n_draw; j += stride) {
    const int64_t i = idx[j];
    if (i < 0 || (size_t)i >= n_src) { atomicOr(flags, 0x80000000u); continue; }  // BoundsError
    out[j] = __ldg(src + i);
  }
}
// the SAME index vector applied to M variables (resampling.jl:80-84): each CTA row
// reads its slice of idx once into registers and reuses it for a group of variables
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_gather_multi(const T* const* __restrict__ srcs, T* const* __restrict__ outs, int M,
               const int64_t* __restrict__ idx, size_t n_draw, size_t n_src, uint32_t* flags) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  const int v0 = blockIdx.y * 8;
  const int v1 = min(v0 + 8, M);
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n_draw; j += stride) {
    const int64_t i = idx[j];
    if (i < 0 || (size_t)i >= n_src) { atomicOr(flags, 0x80000000u); continue; }
    for (int v = v0; v < v1; ++v) outs[v][j] = __ldg(srcs[v] + i);
  }
}

// ---- systematic resampling ----------------------------------------------------------
// Weights are mapped to integers W_i = floor(w_i * 2^s) with s chosen on the host from
// max(w) and n so that sum W_i < 2^62; the inclusive scan is then exact and independent
// of the order of summation (H10), and particle i (0-based) takes ancestor
//   b(i) = min{ b : C_b * n * 2^32 > (i * 2^32 + U) * S },  U = floor(u*2^32), S = C_{n-1}.
__global__ void __launch_bounds__(kThreads)
k_weights_to_fixed(const double* __restrict__ w, unsigned long long* __restrict__ W, size_t n, double scale,
                   uint32_t* flags) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    const double x = w[j];
    if (!(x >= 0.0)) { atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN); W[j] = 0ull; continue; }
    W[j] = (unsigned long long)floor(x * scale);
  }
}
__global__ void __launch_bounds__(kThreads)
k_resample_search(const unsigned long long* __restrict__ Cs, size_t n, unsigned long long U,
                  int64_t* __restrict__ idx_out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  const unsigned long long S = Cs[n - 1];
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const unsigned __int128 rhs = ((unsigned __int128)i * 4294967296ull + U) * (unsigned __int128)S;
    size_t lo = 0, hi = n - 1;  // invariant: answer in [lo, hi]
    while (lo < hi) {
      const size_t mid = lo + ((hi - lo) >> 1);
      const unsigned __int128 lhs = (unsigned __int128)Cs[mid] * (unsigned __int128)n * 4294967296ull;
      if (lhs > rhs) hi = mid; else lo = mid + 1;
    }
    idx_out[i] = (int64_t)lo;
  }
}
__global__ void k_max_abs(const double* __restrict__ w, size_t n, double* out) {
  // single CTA max reduction (n reads, tiny vs the scan); 1024 threads
  __shared__ double sh[32];
  double m = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, w[i]);
  for (int d = 16; d >= 1; d >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int d = 16; d >= 1; d >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, d));
    if (threadIdx.x == 0) *out = m;
  }
}

template <typename U, bool INCLUSIVE> int device_scan(U* a, size_t n, U* grand_total_dev);
template <typename K, bool HAS_VAL> int radix_sort_keys(K* bufA, K* bufB, uint32_t* valA, uint32_t* valB, size_t n);
int read_back(double* host_dst, const double* dev_src, size_t count);
double* moments_result_slot();

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_norminvcdf(int dtype, const void* p_in, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!p_in || !out)) return fail(MCM_ERR_INVALID, "mcm_norminvcdf: null pointer");
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  const int grid = stream_grid(n, 2, 8);
  if (dtype == MCM_F64) k_norminv<double><<<grid, kThreads, 0, c.stream>>>((const double*)p_in, (double*)out, n);
  else if (dtype == MCM_F32) k_norminv<float><<<grid, kThreads, 0, c.stream>>>((const float*)p_in, (float*)out, n);
  else return fail(MCM_ERR_INVALID, "mcm_norminvcdf: bad dtype %d", dtype);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

template <bool NORMAL>
static int systematic_impl(int dtype, void* out, size_t n, double a, double b, const int64_t* perm,
                           uint64_t i0, uint64_t n_total) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_systematic: null pointer");
  if (n_total == 0 || (perm == nullptr && i0 + n > n_total))
    return fail(MCM_ERR_INVALID, "mcm_systematic: index range [%llu,%llu) exceeds n_total %llu",
                (unsigned long long)i0, (unsigned long long)(i0 + n), (unsigned long long)n_total);
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  const int grid = stream_grid(n, 2, 8);
  if (dtype == MCM_F64)
    k_systematic<double, NORMAL><<<grid, kThreads, 0, c.stream>>>((double*)out, n, a, b, perm, i0, n_total);
  else if (dtype == MCM_F32)
    k_systematic<float, NORMAL><<<grid, kThreads, 0, c.stream>>>((float*)out, n, a, b, perm, i0, n_total);
  else return fail(MCM_ERR_INVALID, "mcm_systematic: bad dtype %d", dtype);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

extern "C" int mcm_systematic_normal(int dtype, void* out, size_t n, double mu, double sigma,
                                     const int64_t* perm_dev, uint64_t i0, uint64_t n_total) {
  return systematic_impl<true>(dtype, out, n, mu, sigma, perm_dev, i0, n_total);
}
extern "C" int mcm_systematic_uniform(int dtype, void* out, size_t n, double a, double b,
                                      const int64_t* perm_dev, uint64_t i0, uint64_t n_total) {
  return systematic_impl<false>(dtype, out, n, a, b, perm_dev, i0, n_total);
}

extern "C" int mcm_rand_indices(int64_t* idx_out, size_t n_draw, uint64_t N, uint64_t seed, uint64_t offset) {
  MCM_NEED_INIT();
  if (n_draw && !idx_out) return fail(MCM_ERR_INVALID, "mcm_rand_indices: null pointer");
  if (N == 0) return fail(MCM_ERR_INVALID, "mcm_rand_indices: empty range");
  if (n_draw == 0) return MCM_OK;
  Ctx& c = ctx();
  k_rand_indices<<<stream_grid(n_draw, 2, 8), kThreads, 0, c.stream>>>(idx_out, n_draw, N, seed, offset);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

extern "C" int mcm_randperm(int64_t* perm_out, size_t n, uint64_t seed) {
  MCM_NEED_INIT();
  if (n && !perm_out) return fail(MCM_ERR_INVALID, "mcm_randperm: null pointer");
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  uint64_t *kA = nullptr, *kB = nullptr;
  uint32_t *vA = nullptr, *vB = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&kA, n * 8, c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&kB, n * 8, c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&vA, n * 4, c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&vB, n * 4, c.stream));
  const int grid = stream_grid(n, 2, 8);
  k_rand_keys<<<grid, kThreads, 0, c.stream>>>(kA, vA, n, seed);
  MCM_CHECK_LAUNCH();
  int rc = radix_sort_keys<uint64_t, true>(kA, kB, vA, vB, n);
  if (rc) return rc;
  k_u32_to_i64_b<<<grid, 256, 0, c.stream>>>(vA, perm_out, n);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(kA, c.stream));
  MCM_CUDA(cudaFreeAsync(kB, c.stream));
  MCM_CUDA(cudaFreeAsync(vA, c.stream));
  MCM_CUDA(cudaFreeAsync(vB, c.stream));
  return MCM_OK;
}

extern "C" int mcm_gather(int dtype, const void* src, const int64_t* idx, void* out, size_t n_draw, size_t n_src) {
  MCM_NEED_INIT();
  if (n_draw && (!src || !idx || !out)) return fail(MCM_ERR_INVALID, "mcm_gather: null pointer");
  if (n_draw == 0) return MCM_OK;
  Ctx& c = ctx();
  const int grid = stream_grid(n_draw, 4, 8);
  if (dtype == MCM_F64) k_gather<double><<<grid, kThreads, 0, c.stream>>>((const double*)src, idx, (double*)out, n_draw, n_src, c.d_flags);
  else if (dtype == MCM_F32) k_gather<float><<<grid, kThreads, 0, c.stream>>>((const float*)src, idx, (float*)out, n_draw, n_src, c.d_flags);
  else return fail(MCM_ERR_INVALID, "mcm_gather: bad dtype %d", dtype);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

extern "C" int mcm_gather_multi(int dtype, const void* const* srcs, void* const* outs, int M,
                                const int64_t* idx, size_t n_draw, size_t n_src) {
  MCM_NEED_INIT();
  if (M < 1 || !srcs || !outs || (n_draw && !idx)) return fail(MCM_ERR_INVALID, "mcm_gather_multi: bad arguments");
  if (n_draw == 0) return MCM_OK;
  Ctx& c = ctx();
  void** d_tab = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_tab, 2 * (size_t)M * sizeof(void*), c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_tab, srcs, (size_t)M * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_tab + M, outs, (size_t)M * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
  // grid.y walks groups of 8 variables; grid.x covers the draws (grid-stride), sized so
  // that the whole grid is about 8 CTAs per SM
  const int gy = (M + 7) / 8;
  int gx = stream_grid(n_draw, 1, 8);
  int cap = (8 * c.sm_count + gy - 1) / gy;
  if (cap < 4) cap = 4;
  if (gx > cap) gx = cap;
  dim3 grid(gx, gy);
  if (dtype == MCM_F64)
    k_gather_multi<double><<<grid, kThreads, 0, c.stream>>>((const double* const*)d_tab, (double* const*)(d_tab + M), M, idx, n_draw, n_src, c.d_flags);
  else if (dtype == MCM_F32)
    k_gather_multi<float><<<grid, kThreads, 0, c.stream>>>((const float* const*)d_tab, (float* const*)(d_tab + M), M, idx, n_draw, n_src, c.d_flags);
  else return fail(MCM_ERR_INVALID, "mcm_gather_multi: bad dtype %d", dtype);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(d_tab, c.stream));
  return MCM_OK;
}

extern "C" int mcm_resample_systematic(const double* w, size_t n, double u, int64_t* idx_out) {
  MCM_NEED_INIT();
  if (!w || !idx_out || n == 0) return fail(MCM_ERR_INVALID, "mcm_resample_systematic: bad arguments");
  if (!(u >= 0.0 && u < 1.0)) return fail(MCM_ERR_INVALID, "mcm_resample_systematic: u must be in [0,1)");
  if (n > (1ull << 30)) return fail(MCM_ERR_UNSUPPORTED, "mcm_resample_systematic: n must be <= 2^30");
  Ctx& c = ctx();
  double* d_max = moments_result_slot() + 10;
  k_max_abs<<<1, 1024, 0, c.stream>>>(w, n, d_max);
  MCM_CHECK_LAUNCH();
  double wmax;
  int rc = read_back(&wmax, d_max, 1);
  if (rc) return rc;
  if (!(wmax > 0.0) || !isfinite(wmax)) return fail(MCM_ERR_INVALID, "mcm_resample_systematic: weights must be finite with a positive maximum");
  // scale = 2^s with wmax*2^s < 2^(62 - ceil(log2 n))  ->  sum < 2^62
  int e;
  frexp(wmax, &e);  // wmax < 2^e
  int lg = 0;
  while ((1ull << lg) < n) ++lg;
  const int s = 62 - lg - e;
  const double scale = ldexp(1.0, s);
  unsigned long long* W = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&W, n * sizeof(unsigned long long), c.stream));
  const int grid = stream_grid(n, 2, 8);
  k_weights_to_fixed<<<grid, kThreads, 0, c.stream>>>(w, W, n, scale, c.d_flags);
  MCM_CHECK_LAUNCH();
  rc = device_scan<unsigned long long, true>(W, n, nullptr);
  if (rc) return rc;
  const unsigned long long U = (unsigned long long)floor(u * 4294967296.0);
  k_resample_search<<<grid, kThreads, 0, c.stream>>>(W, n, U, idx_out);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(W, c.stream));
  return MCM_OK;
}
