// sampling.cu — systematic_sample (src/sampling.jl:6-17), bootstrap
// (src/resampling.jl:76-84) and systematic resampling (the dead reference code at
// src/resampling.jl:45-67, restated with an exact integer scan).
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_norminv.cuh"
#include "mcm_dist.cuh"

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
// iid uniforms in (0,1) and standard normals from the same counters: u = (k + 1/2) 2^-53 with k the top 53 bits
// of a Philox draw (never 0 or 1), z = Phi^-1(u).  rand(rng, Normal(), N) stand-in for the iid constructors
// (types.jl:67-82 with systematic=false, types.jl:92-100) and optimize's perturb (optimize.jl:3-6).
template <typename T, bool NORMAL>
__global__ void __launch_bounds__(kThreads)
k_rand_real(T* __restrict__ out, size_t n, unsigned long long seed, unsigned long long offset, double a, double b) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    uint32_t r[4];
    philox4x32(offset + j, seed, r);
    const unsigned long long k = ((((unsigned long long)r[1] << 32) | r[0]) >> 11);
    const double u = ((double)k + 0.5) * 0x1p-53;
    double v;
    if constexpr (NORMAL) v = a + b * norminv::norminvcdf(u); else v = a + (b - a) * u;
    out[j] = (T)v;
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

// ---- keyed index bijection ---------------------------------------------------------
// randperm(rng, N) (sampling.jl:11) without materialising or sorting anything: a keyed
// pseudo-random PERMUTATION of [0, n_total) evaluated per index — a 6-round alternating
// Feistel network on ceil(log2 n_total) bits (left half floor(b/2) bits, right half the rest),
// cycle-walked back into range (domain < 2 n_total, so < 2 evaluations on average).  Every rank of
// a sharded sample evaluates its own index range with no communication.
// The round function is a 32-bit multiply-xorshift mixer; round keys come from splitmix64 of
// (seed, column).  oracle/mcm_oracle.c restates it bit for bit.  Rounds: 2 fail simple correlation
// tests, 4 pass them (lag correlations, index correlation, pair chi-square, 3-column moments at
// n = 1e3..1e6); 6 are used for margin (8 cost 16.3 ms per 1e9 samples).
// 4 rounds (6 until the last session of round 2): with a full-avalanche round function 4 rounds are the classical
// strong-PRP count, and the rank correlations between columns, with the identity and between neighbours, and the
// 64 x 64 chi-square of the joint position tables are the same as with 6 (tests/test_host_logic.py, scripts output
// in DESIGN.md); 2 rounds fail all of them.  The batched build of C5 is bound by exactly this integer work.
constexpr int kFeistelRounds = 4;
struct FeistelKey {
  uint32_t k[8];
  uint32_t lbits, rbits;
};
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ FeistelKey feistel_key(uint64_t seed, uint64_t column, uint64_t n_total) {
  FeistelKey fk;
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (column + 1));
#pragma unroll
  for (int r = 0; r < 8; ++r) fk.k[r] = (uint32_t)(splitmix64(s) >> 32);
  uint32_t bits = 2;
  while (bits < 64 && (1ull << bits) < n_total) ++bits;
  fk.lbits = bits / 2;
  fk.rbits = bits - fk.lbits;
  return fk;
}
__host__ __device__ __forceinline__ uint32_t feistel_mix(uint32_t v, uint32_t k) {
  uint32_t h = v + k;
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint64_t feistel_permute(uint64_t x, const FeistelKey& fk, uint64_t n_total) {
  const uint32_t lmask = fk.lbits >= 32 ? 0xffffffffu : ((1u << fk.lbits) - 1u);
  const uint32_t rmask = fk.rbits >= 32 ? 0xffffffffu : ((1u << fk.rbits) - 1u);
  do {
    uint32_t L = (uint32_t)(x >> fk.rbits) & lmask, R = (uint32_t)x & rmask;
#pragma unroll
    for (int r = 0; r < kFeistelRounds; r += 2) {
      L ^= feistel_mix(R, fk.k[r]) & lmask;
      R ^= feistel_mix(L, fk.k[r + 1]) & rmask;
    }
    x = ((uint64_t)L << fk.rbits) | R;
  } while (x >= n_total);
  return x;
}

// inverse of feistel_permute: the rounds backwards, walked along the same cycle until the value is in range
__host__ __device__ __forceinline__ uint64_t feistel_inverse(uint64_t x, const FeistelKey& fk, uint64_t n_total) {
  const uint32_t lmask = fk.lbits >= 32 ? 0xffffffffu : ((1u << fk.lbits) - 1u);
  const uint32_t rmask = fk.rbits >= 32 ? 0xffffffffu : ((1u << fk.rbits) - 1u);
  do {
    uint32_t L = (uint32_t)(x >> fk.rbits) & lmask, R = (uint32_t)x & rmask;
#pragma unroll
    for (int r = kFeistelRounds - 2; r >= 0; r -= 2) {
      R ^= feistel_mix(L, fk.k[r + 1]) & rmask;
      L ^= feistel_mix(R, fk.k[r]) & lmask;
    }
    x = ((uint64_t)L << fk.rbits) | R;
  } while (x >= n_total);
  return x;
}

__global__ void __launch_bounds__(kThreads)
k_index_bijection(int64_t* __restrict__ out, size_t n, FeistelKey fk, unsigned long long i0, unsigned long long n_total) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride)
    out[j] = (int64_t)feistel_permute(i0 + j, fk, n_total);
}

// M systematic columns in one launch (grid.y = column): out_v[j] = a_v + b_v * Phi^-1(grid[pi_v(i0+j)]).
// ab: [2M] doubles on device (a then b) or nullptr for (0, 1) / Uniform(0,1).
template <typename T, bool NORMAL>
__global__ void __launch_bounds__(kThreads)
k_systematic_batch(T* const* __restrict__ outs, size_t n, const double* __restrict__ ab, int M, int permute,
                   unsigned long long seed, unsigned long long i0, unsigned long long n_total) {
  const int v = blockIdx.y;
  T* __restrict__ out = outs[v];
  const double a = ab ? ab[v] : 0.0, b = ab ? ab[M + v] : 1.0;
  const FeistelKey fk = feistel_key(seed, (uint64_t)v, n_total);
  const double two_n = 2.0 * (double)n_total;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    const unsigned long long i = permute ? feistel_permute(i0 + j, fk, n_total) : i0 + j;
    const double p = (double)(2ull * i + 1ull) / two_n;
    double val;
    if constexpr (NORMAL) val = a + b * norminv::norminvcdf(p);
    else val = a + (b - a) * p;
    __stcs(out + j, (T)val);
  }
}

// Whole columns (i0 == 0, n == n_total): SCATTER form.  Thread i evaluates grid point i — consecutive lanes take
// consecutive probabilities, so a warp is either entirely in the central region of AS 241 or entirely in a tail (the
// gather form, with permuted probabilities, runs both branches in almost every warp: 15 % of the lanes need the
// log/sqrt tail) — and stores the value at j = pi^-1(i).  grid.x covers a whole column and is the fast grid index, so
// only one or two columns (8 MB each at n = 1e6) are being written at any time: the scattered 8-byte stores merge in
// L2 and reach HBM as full lines.  Same values at the same positions as the gather form.
constexpr int SB_IPT = 4;
template <typename T, bool NORMAL>
__global__ void __launch_bounds__(kThreads)
k_systematic_batch_scatter(T* const* __restrict__ outs, size_t n, const double* __restrict__ ab, int M, int permute,
                           unsigned long long seed) {
  const int v = blockIdx.y;
  T* __restrict__ out = outs[v];
  const double a = ab ? ab[v] : 0.0, b = ab ? ab[M + v] : 1.0;
  const FeistelKey fk = feistel_key(seed, (uint64_t)v, n);
  const double two_n = 2.0 * (double)n;
#pragma unroll
  for (int u = 0; u < SB_IPT; ++u) {
    const size_t i = (size_t)blockIdx.x * (kThreads * SB_IPT) + (size_t)u * kThreads + threadIdx.x;
    if (i < n) {
      const double p = (double)(2ull * i + 1ull) / two_n;
      double val;
      if constexpr (NORMAL) val = a + b * norminv::norminvcdf(p);
      else val = a + (b - a) * p;
      const size_t j = permute ? (size_t)feistel_inverse(i, fk, n) : i;
      out[j] = (T)val;
    }
  }
}

// Every column of a batch is a permutation of the SAME grid z_i = Phi^-1((2i+1)/(2n)) (src/sampling.jl:8-10), so the
// grid is evaluated once (k_sys_grid, n values) and the columns read it back through L2 (8 MB at n = 1e6, read by all
// M columns): C5's 1024 x 1e6 build drops from 1e9 inverse-CDF evaluations to 1e6.  Same values: a + b * z_i with
// the same two roundings.
__global__ void __launch_bounds__(kThreads)
k_sys_grid(double* __restrict__ z, size_t n) {
  const double two_n = 2.0 * (double)n;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    z[i] = norminv::norminvcdf((double)(2ull * i + 1ull) / two_n);
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_systematic_batch_scatter_tab(T* const* __restrict__ outs, size_t n, const double* __restrict__ ab, int M,
                               unsigned long long seed, const double* __restrict__ z) {
  const int v = blockIdx.y;
  T* __restrict__ out = outs[v];
  const double a = ab ? ab[v] : 0.0, b = ab ? ab[M + v] : 1.0;
  const FeistelKey fk = feistel_key(seed, (uint64_t)v, n);
#pragma unroll
  for (int u = 0; u < SB_IPT; ++u) {
    const size_t i = (size_t)blockIdx.x * (kThreads * SB_IPT) + (size_t)u * kThreads + threadIdx.x;
    if (i < n) {
      const double val = a + b * z[i];
      out[(size_t)feistel_inverse(i, fk, n)] = (T)val;
    }
  }
}

// gather form of the same: coalesced streaming stores, the table read at the permuted index (L2-resident)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_systematic_batch_gather_tab(T* const* __restrict__ outs, size_t n, const double* __restrict__ ab, int M, unsigned long long seed,
                              unsigned long long i0, unsigned long long n_total, const double* __restrict__ z) {
  const int v = blockIdx.y;
  T* __restrict__ out = outs[v];
  const double a = ab ? ab[v] : 0.0, b = ab ? ab[M + v] : 1.0;
  const FeistelKey fk = feistel_key(seed, (uint64_t)v, n_total);
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    const unsigned long long i = feistel_permute(i0 + j, fk, n_total);
    __stcs(out + j, (T)(a + b * z[i]));
  }
}

// ---- any univariate distribution (mcm_dist.cuh) ----------------------------------------------------------------
// MODE 0: systematic grid (src/sampling.jl:8-10): p_i = (2i+1)/(2N) and q_i = 1 - p_i = (2(N-1-i)+1)/(2N), both
//         correctly rounded quotients of exact integers, so the upper tail is as accurate as the lower one;
// MODE 1: iid draws through the inverse CDF of a Philox uniform (types.jl:71,79 with systematic = false);
// MODE 2: quantile.(d, p) on a caller-supplied probability vector.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads)
k_dist(T* __restrict__ out, size_t n, int kind, dist::Params dp, const int64_t* __restrict__ perm, const T* __restrict__ pin,
       unsigned long long i0, unsigned long long n_total, unsigned long long seed, unsigned long long offset, uint32_t* flags) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  const double two_n = 2.0 * (double)n_total;
  bool bad = false;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
    double p, q;
    if constexpr (MODE == 0) {
      const unsigned long long i = perm ? (unsigned long long)perm[j] : i0 + j;
      p = (double)(2ull * i + 1ull) / two_n;
      q = (double)(2ull * (n_total - 1ull - i) + 1ull) / two_n;
    } else if constexpr (MODE == 1) {
      uint32_t r[4];
      philox4x32(offset + j, seed, r);
      const unsigned long long k = ((((unsigned long long)r[1] << 32) | r[0]) >> 11);
      p = ((double)k + 0.5) * 0x1p-53;
      q = ((double)(9007199254740991ull - k) + 0.5) * 0x1p-53;          // 1 - p, exactly
    } else {
      p = (double)pin[j];
      q = 1.0 - p;
    }
    const double v = dist::quantile(kind, dp, p, q);
    bad |= (v != v) && (p == p);
    out[j] = (T)v;
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
}

// ---- gather ----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_gather(const T* __restrict__ src, const int64_t* __restrict__ idx, T* __restrict__ out, size_t n_draw,
         size_t n_src, uint32_t* flags) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n_draw; j += stride) {
    const int64_t i = idx[j];
    if (i < 0 || (size_t)i >= n_src) {       // BoundsError in the reference: sticky flag, and a NaN so it cannot pass silently
      atomicOr(flags, 0x80000000u);
      out[j] = (T)__builtin_nan("");
      continue;
    }
    out[j] = __ldg(src + i);
  }
}
// the SAME index vector applied to M variables (resampling.jl:80-84).  grid.y walks the
// variables four at a time, grid.x covers ALL draws (no grid-stride), and CTAs are scheduled
// x-fastest: at any moment the resident CTAs work on one or two groups of variables, whose
// source arrays (4 x 8 MB at N = 1e6) stay in the 126 MB L2 — the random 8-byte reads then
// hit L2 and HBM sees each source once, sequentially (first version: every group in flight
// at once, 8 GB working set, 686 GB/s).
constexpr int GM_VARS = 4;
constexpr int GM_IPT = 4;
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_gather_multi(const T* const* __restrict__ srcs, T* const* __restrict__ outs, int M,
               const int64_t* __restrict__ idx, size_t n_draw, size_t n_src, uint32_t* flags) {
  const int v0 = blockIdx.y * GM_VARS;
  const size_t base = (size_t)blockIdx.x * (kThreads * GM_IPT) + threadIdx.x;
  int64_t ii[GM_IPT];
  bool ok[GM_IPT];
#pragma unroll
  for (int u = 0; u < GM_IPT; ++u) {
    const size_t j = base + (size_t)u * kThreads;
    ii[u] = j < n_draw ? __ldg(idx + j) : 0;
    ok[u] = j < n_draw;
    if (ok[u] && (ii[u] < 0 || (size_t)ii[u] >= n_src)) { atomicOr(flags, 0x80000000u); ok[u] = false; }  // BoundsError
  }
#pragma unroll
  for (int dv = 0; dv < GM_VARS; ++dv) {
    const int v = v0 + dv;
    if (v >= M) break;
    const T* __restrict__ src = srcs[v];
    T* __restrict__ out = outs[v];
    T val[GM_IPT];
#pragma unroll
    for (int u = 0; u < GM_IPT; ++u) val[u] = ok[u] ? __ldg(src + ii[u]) : (T)__builtin_nan("");
#pragma unroll
    for (int u = 0; u < GM_IPT; ++u)
      if (base + (size_t)u * kThreads < n_draw) __stcs(out + base + (size_t)u * kThreads, val[u]);
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
// Expansion form of the same rule (first version: one 27-step binary search per output over the 800 MB scan,
// 50 ms at n = 1e8).  Source b owns the outputs [E(C_{b-1}), E(C_b)) with
//   E(C) = #{ i in [0,n) : (i 2^32 + U) S < C n 2^32 } = min(n, max(0, floor((ceil(C n 2^32 / S) - 1 - U) / 2^32) + 1)),
// one exact 128-bit division per source, streaming reads of the scan and (almost) streaming writes of the indices.
__device__ __forceinline__ unsigned long long resample_end(unsigned long long C, unsigned long long n, unsigned long long U,
                                                           unsigned long long S) {
  const unsigned __int128 X = (unsigned __int128)C * ((unsigned __int128)n << 32);
  const unsigned __int128 Q = (X + (S - 1)) / S;          // ceil; < 2^63 because C <= S
  if (Q < (unsigned __int128)U + 1) return 0ull;
  const unsigned long long e = (unsigned long long)((Q - 1 - U) >> 32) + 1ull;
  return e < n ? e : n;
}
constexpr unsigned RS_SMALL = 8;          // runs up to this length are written by the owning thread
constexpr unsigned RS_CHUNK = 1u << 16;   // longer-than-a-warp-job runs are cut into chunks for k_resample_fill
struct RsBig { unsigned long long start; unsigned len; unsigned src; };

// Where the indices go: this rank's array, or (sharded) the index windows of the ranks that own the
// output positions — plain P2P stores over NVLink, runs are contiguous so they coalesce.
struct RsOutLocal {
  int64_t* p;
  __device__ __forceinline__ void put(unsigned long long k, long long v) const { p[k] = v; }
};
struct RsOutPeer {
  int64_t* win[16];
  unsigned long long per;          // output positions per rank
  __device__ __forceinline__ void put(unsigned long long k, long long v) const {
    const unsigned d = (unsigned)(k / per);
    win[d][k - (unsigned long long)d * per] = v;
  }
};
// Sharded form: this rank holds sources [src0, src0 + n) of the global weight vector; C0 = sum of the fixed-point weights
// of all lower ranks, S = global total, n_out = global number of outputs.  Single rank: src0 = C0 = 0, n_out = n.
template <class Out>
__global__ void __launch_bounds__(kThreads)
k_resample_expand(const unsigned long long* __restrict__ Cs, size_t n, unsigned long long U, const __grid_constant__ Out out,
                  RsBig* __restrict__ big, unsigned* __restrict__ nbig, unsigned big_cap, unsigned long long C0,
                  unsigned long long S, unsigned long long n_out, unsigned long long src0) {
  const int lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t rounds = (n + stride - 1) / stride;   // warp-uniform trip count (shuffles inside)
  for (size_t r = 0; r < rounds; ++r) {
    const size_t b = r * stride + (size_t)blockIdx.x * kThreads + threadIdx.x;
    const bool live = b < n;
    const unsigned long long eb = live ? resample_end(C0 + Cs[b], n_out, U, S) : n_out;
    unsigned long long ea = __shfl_up_sync(0xffffffffu, eb, 1);
    if (lane == 0) ea = live ? resample_end(C0 + (b > 0 ? Cs[b - 1] : 0ull), n_out, U, S) : 0ull;
    const unsigned long long cnt = live ? eb - ea : 0ull;
    if (cnt && cnt <= RS_SMALL) {
      for (unsigned k = 0; k < (unsigned)cnt; ++k) out.put(ea + k, (long long)(src0 + b));
    }
    // medium runs: the warp fills them together, coalesced
    unsigned med = __ballot_sync(0xffffffffu, cnt > RS_SMALL && cnt <= RS_CHUNK);
    while (med) {
      const int src = __ffs(med) - 1;
      med &= med - 1;
      const unsigned long long s0 = __shfl_sync(0xffffffffu, ea, src);
      const unsigned len = (unsigned)__shfl_sync(0xffffffffu, cnt, src);
      const long long who = (long long)(src0 + __shfl_sync(0xffffffffu, (unsigned long long)b, src));
      for (unsigned k = lane; k < len; k += 32) out.put(s0 + k, who);
    }
    // long runs (a few dominant weights): queued in chunks for a grid-wide fill
    if (cnt > RS_CHUNK) {
      const unsigned nch = (unsigned)((cnt + RS_CHUNK - 1) / RS_CHUNK);
      const unsigned at = atomicAdd(nbig, nch);
      for (unsigned q = 0; q < nch; ++q) {
        if (at + q < big_cap) {
          const unsigned long long off = (unsigned long long)q * RS_CHUNK;
          big[at + q].start = ea + off;
          big[at + q].len = (unsigned)min((unsigned long long)RS_CHUNK, cnt - off);
          big[at + q].src = (unsigned)b;
        }
      }
    }
  }
}
template <class Out>
__global__ void __launch_bounds__(kThreads)
k_resample_fill(const RsBig* __restrict__ big, const unsigned* __restrict__ nbig, const __grid_constant__ Out out,
                unsigned long long src0) {
  const unsigned total = *nbig;
  for (unsigned q = blockIdx.x; q < total; q += gridDim.x) {
    const RsBig e = big[q];
    for (unsigned k = threadIdx.x; k < e.len; k += kThreads) out.put(e.start + k, (long long)(src0 + e.src));
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
int moments_device(int dtype, const void* in, size_t n, double* d_out6);

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

extern "C" int mcm_index_bijection(int64_t* out, size_t n, uint64_t seed, uint64_t column, uint64_t i0, uint64_t n_total) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_index_bijection: null pointer");
  if (n_total == 0 || i0 + n > n_total) return fail(MCM_ERR_INVALID, "mcm_index_bijection: index range exceeds n_total");
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  k_index_bijection<<<stream_grid(n, 2, 8), kThreads, 0, c.stream>>>(out, n, feistel_key(seed, column, n_total), i0, n_total);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

static int g_sys_scatter = 1;   // A/B: 0 = always the gather form
extern "C" int mcm_systematic_use_scatter(int on) { g_sys_scatter = on; return MCM_OK; }

extern "C" int mcm_systematic_batch(int dtype, int kind, void* const* outs, int M, size_t n, const double* a_host,
                                    const double* b_host, int permute, uint64_t seed, uint64_t i0, uint64_t n_total) {
  MCM_NEED_INIT();
  if (M < 1 || !outs) return fail(MCM_ERR_INVALID, "mcm_systematic_batch: bad arguments");
  if (M > 65535) return fail(MCM_ERR_UNSUPPORTED, "mcm_systematic_batch: at most 65535 columns per call");
  if ((a_host == nullptr) != (b_host == nullptr)) return fail(MCM_ERR_INVALID, "mcm_systematic_batch: pass both parameter arrays or neither");
  if (kind != 0 && kind != 1) return fail(MCM_ERR_INVALID, "mcm_systematic_batch: kind must be 0 (Normal) or 1 (Uniform)");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_systematic_batch: bad dtype %d", dtype);
  if (n_total == 0 || i0 + n > n_total) return fail(MCM_ERR_INVALID, "mcm_systematic_batch: index range exceeds n_total");
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  std::vector<unsigned char> host((size_t)M * sizeof(void*) + (a_host ? 2 * (size_t)M * sizeof(double) : 0));
  memcpy(host.data(), outs, (size_t)M * sizeof(void*));
  if (a_host) {
    memcpy(host.data() + (size_t)M * sizeof(void*), a_host, (size_t)M * sizeof(double));
    memcpy(host.data() + (size_t)M * sizeof(void*) + (size_t)M * sizeof(double), b_host, (size_t)M * sizeof(double));
  }
  unsigned char* d_tab = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_tab, host.size(), c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_tab, host.data(), host.size(), cudaMemcpyHostToDevice, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));  // host staging vector dies at return
  const double* d_ab = a_host ? (const double*)(d_tab + (size_t)M * sizeof(void*)) : nullptr;
  const size_t gx_full = (n + (size_t)kThreads * SB_IPT - 1) / ((size_t)kThreads * SB_IPT);
  if (g_sys_scatter && permute && i0 == 0 && n == n_total && n * (dtype == MCM_F64 ? 8 : 4) <= (32u << 20) && gx_full <= 0x7fffffffull) {
    // whole columns that fit in L2 several times over: scatter form (coherent Phi^-1 branches)
    dim3 grid((unsigned)gx_full, (unsigned)M);
    if (kind == 0 && M >= 2) {
      double* z = nullptr;
      MCM_CUDA(cudaMallocAsync((void**)&z, n * sizeof(double), c.stream));
      k_sys_grid<<<stream_grid(n, 1, 8), kThreads, 0, c.stream>>>(z, n);
      MCM_CHECK_LAUNCH();
      if (g_sys_scatter != 3) {        // measured at C5 (1024 x 1e6): gather 7.65 ms, scatter 9.8 ms (mcm_systematic_use_scatter(3))
        int gx = stream_grid(n, 2, 8);
        const int cap = std::max(1, (8 * c.sm_count + M - 1) / M);
        if (gx > cap) gx = cap;
        dim3 gg((unsigned)gx, (unsigned)M);
        if (dtype == MCM_F64) k_systematic_batch_gather_tab<double><<<gg, kThreads, 0, c.stream>>>((double* const*)d_tab, n, d_ab, M, seed, 0, n, z);
        else k_systematic_batch_gather_tab<float><<<gg, kThreads, 0, c.stream>>>((float* const*)d_tab, n, d_ab, M, seed, 0, n, z);
      } else if (dtype == MCM_F64) k_systematic_batch_scatter_tab<double><<<grid, kThreads, 0, c.stream>>>((double* const*)d_tab, n, d_ab, M, seed, z);
      else k_systematic_batch_scatter_tab<float><<<grid, kThreads, 0, c.stream>>>((float* const*)d_tab, n, d_ab, M, seed, z);
      MCM_CHECK_LAUNCH();
      MCM_CUDA(cudaFreeAsync(z, c.stream));
      MCM_CUDA(cudaFreeAsync(d_tab, c.stream));
      return MCM_OK;
    }
#define MCM_SYS_SCATTER(T, NORMAL) \
  k_systematic_batch_scatter<T, NORMAL><<<grid, kThreads, 0, c.stream>>>((T* const*)d_tab, n, d_ab, M, permute, seed)
    if (dtype == MCM_F64) { if (kind == 0) MCM_SYS_SCATTER(double, true); else MCM_SYS_SCATTER(double, false); }
    else { if (kind == 0) MCM_SYS_SCATTER(float, true); else MCM_SYS_SCATTER(float, false); }
#undef MCM_SYS_SCATTER
  } else {
    int gx = stream_grid(n, 2, 8);
    const int cap = std::max(1, (8 * c.sm_count + M - 1) / M);
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)M);
#define MCM_SYS_BATCH(T, NORMAL) \
  k_systematic_batch<T, NORMAL><<<grid, kThreads, 0, c.stream>>>((T* const*)d_tab, n, d_ab, M, permute, seed, i0, n_total)
    if (dtype == MCM_F64) { if (kind == 0) MCM_SYS_BATCH(double, true); else MCM_SYS_BATCH(double, false); }
    else { if (kind == 0) MCM_SYS_BATCH(float, true); else MCM_SYS_BATCH(float, false); }
#undef MCM_SYS_BATCH
  }
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(d_tab, c.stream));
  return MCM_OK;
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

extern "C" int mcm_rand_real(int dtype, int kind, void* out, size_t n, double a, double b, uint64_t seed, uint64_t offset) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_rand_real: null pointer");
  if (kind != 0 && kind != 1) return fail(MCM_ERR_INVALID, "mcm_rand_real: kind must be 0 (Normal) or 1 (Uniform)");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_rand_real: bad dtype %d", dtype);
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  const int grid = stream_grid(n, 2, 8);
  if (dtype == MCM_F64) {
    if (kind == 0) k_rand_real<double, true><<<grid, kThreads, 0, c.stream>>>((double*)out, n, seed, offset, a, b);
    else k_rand_real<double, false><<<grid, kThreads, 0, c.stream>>>((double*)out, n, seed, offset, a, b);
  } else {
    if (kind == 0) k_rand_real<float, true><<<grid, kThreads, 0, c.stream>>>((float*)out, n, seed, offset, a, b);
    else k_rand_real<float, false><<<grid, kThreads, 0, c.stream>>>((float*)out, n, seed, offset, a, b);
  }
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
  const unsigned gy = (unsigned)((M + GM_VARS - 1) / GM_VARS);
  const size_t gx = (n_draw + (size_t)kThreads * GM_IPT - 1) / ((size_t)kThreads * GM_IPT);
  if (gy > 65535u || gx > 0x7fffffffull) return fail(MCM_ERR_UNSUPPORTED, "mcm_gather_multi: problem too large for one launch");
  dim3 grid((unsigned)gx, gy);
  if (dtype == MCM_F64)
    k_gather_multi<double><<<grid, kThreads, 0, c.stream>>>((const double* const*)d_tab, (double* const*)(d_tab + M), M, idx, n_draw, n_src, c.d_flags);
  else if (dtype == MCM_F32)
    k_gather_multi<float><<<grid, kThreads, 0, c.stream>>>((const float* const*)d_tab, (float* const*)(d_tab + M), M, idx, n_draw, n_src, c.d_flags);
  else return fail(MCM_ERR_INVALID, "mcm_gather_multi: bad dtype %d", dtype);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(d_tab, c.stream));
  return MCM_OK;
}

static int g_resample_search = 0;   // A/B: 1 = one binary search per output (first version)
extern "C" int mcm_resample_use_search(int on) { g_resample_search = on; return MCM_OK; }

extern "C" int mcm_resample_systematic(const double* w, size_t n, double u, int64_t* idx_out) {
  MCM_NEED_INIT();
  if (!w || !idx_out || n == 0) return fail(MCM_ERR_INVALID, "mcm_resample_systematic: bad arguments");
  if (!(u >= 0.0 && u < 1.0)) return fail(MCM_ERR_INVALID, "mcm_resample_systematic: u must be in [0,1)");
  if (n > (1ull << 30)) return fail(MCM_ERR_UNSUPPORTED, "mcm_resample_systematic: n must be <= 2^30");
  Ctx& c = ctx();
  // largest weight: the grid-wide one-pass moments kernel (first version: a single-CTA max loop, 45 of the 50 ms)
  double m6[6];
  int rc = moments_device(MCM_F64, w, n, moments_result_slot());
  if (rc) return rc;
  rc = read_back(m6, moments_result_slot(), 6);
  if (rc) return rc;
  const double wmax = m6[4];
  if (m6[3] < 0.0) return fail(MCM_ERR_INVALID, "mcm_resample_systematic: weights must be non-negative");
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
  if (g_resample_search) {
    k_resample_search<<<grid, kThreads, 0, c.stream>>>(W, n, U, idx_out);
    MCM_CHECK_LAUNCH();
  } else {
    // every output belongs to exactly one run; runs longer than RS_CHUNK are split, so at most n / RS_CHUNK + n_runs
    // chunks exist and n / RS_CHUNK + 2 * (number of long runs) <= 3 n / RS_CHUNK + 1 bounds the queue
    const unsigned big_cap = (unsigned)(3 * (n / RS_CHUNK) + 64);
    RsBig* d_big = nullptr;
    unsigned* d_nbig = nullptr;
    MCM_CUDA(cudaMallocAsync((void**)&d_big, (size_t)big_cap * sizeof(RsBig) + 16, c.stream));
    d_nbig = reinterpret_cast<unsigned*>(d_big + big_cap);
    MCM_CUDA(cudaMemsetAsync(d_nbig, 0, sizeof(unsigned), c.stream));
    unsigned long long S = 0;
    MCM_CUDA(cudaMemcpyAsync(&S, W + (n - 1), sizeof(S), cudaMemcpyDeviceToHost, c.stream));
    MCM_CUDA(cudaStreamSynchronize(c.stream));
    const RsOutLocal out{idx_out};
    k_resample_expand<RsOutLocal><<<grid, kThreads, 0, c.stream>>>(W, n, U, out, d_big, d_nbig, big_cap, 0ull, S, n, 0ull);
    MCM_CHECK_LAUNCH();
    k_resample_fill<RsOutLocal><<<c.sm_count * 4, kThreads, 0, c.stream>>>(d_big, d_nbig, out, 0ull);
    MCM_CHECK_LAUNCH();
    MCM_CUDA(cudaFreeAsync(d_big, c.stream));
  }
  MCM_CUDA(cudaFreeAsync(W, c.stream));
  return MCM_OK;
}

namespace mcm {
int window_lookup(int id, void** peers16, int* world, int* rank, size_t* bytes);
int comm_world(int* nranks, int* rank);
int comm_allgather_bytes(const void* send, void* recv, size_t bytes_per_rank);
}

// Systematic resampling of a particle-sharded weight vector: this rank holds w[global_lo ..
// global_lo + n_local) of n_global weights.  The rule is mcm_resample_systematic's (same fixed-point scale from the
// GLOBAL maximum and n_global, so the indices are identical to the single-process ones); only 2 doubles and 1 integer
// per rank cross NVLink for the scale and the scan offsets.  Every source writes its run of output positions into the
// index window of whichever rank owns them (`window`: per_out int64 per rank, output k lives on rank k / per_out).
// Collective; bracket the later read of the window with mcm_comm_fence().
extern "C" int mcm_resample_systematic_sharded(const double* w_local, size_t n_local, uint64_t n_global, uint64_t global_lo,
                                               uint64_t n_out, double u, int window, uint64_t per_out) {
  MCM_NEED_INIT();
  if ((n_local && !w_local) || n_global == 0 || n_out == 0 || per_out == 0)
    return fail(MCM_ERR_INVALID, "mcm_resample_systematic_sharded: bad arguments");
  if (!(u >= 0.0 && u < 1.0)) return fail(MCM_ERR_INVALID, "mcm_resample_systematic_sharded: u must be in [0,1)");
  if (n_global > (1ull << 30) || n_out > (1ull << 30)) return fail(MCM_ERR_UNSUPPORTED, "mcm_resample_systematic_sharded: sizes must be <= 2^30");
  Ctx& c = ctx();
  int world, rank;
  comm_world(&world, &rank);
  void* peers[16];
  int wworld, wrank;
  size_t wbytes;
  int rc = window_lookup(window, peers, &wworld, &wrank, &wbytes);
  if (rc) return rc;
  if (wworld != world || per_out * sizeof(int64_t) > wbytes || (n_out + per_out - 1) / per_out > (uint64_t)world)
    return fail(MCM_ERR_INVALID, "mcm_resample_systematic_sharded: window does not fit the output layout");
  // 1. global extrema of the weights: local one-pass moments, 16 bytes per rank all-gathered
  double m6[6] = {0, 0, 0, __builtin_huge_val(), -__builtin_huge_val(), 0};
  if (n_local) {
    rc = moments_device(MCM_F64, w_local, n_local, moments_result_slot());
    if (rc) return rc;
    rc = read_back(m6, moments_result_slot(), 6);
    if (rc) return rc;
  }
  unsigned long long* d_x = nullptr;     // exchange buffer: [mine(2)] [all(2*world)]
  MCM_CUDA(cudaMallocAsync((void**)&d_x, sizeof(double) * 2 * (size_t)(world + 1), c.stream));
  std::vector<double> hx(2 * (size_t)world);
  const double mine[2] = {m6[3], m6[4]};
  MCM_CUDA(cudaMemcpyAsync(d_x, mine, 16, cudaMemcpyHostToDevice, c.stream));
  rc = comm_allgather_bytes(d_x, d_x + 2, 16);
  if (rc) return rc;
  MCM_CUDA(cudaMemcpyAsync(hx.data(), d_x + 2, 16 * (size_t)world, cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  double wmin = __builtin_huge_val(), wmax = -__builtin_huge_val();
  for (int r = 0; r < world; ++r) {
    if (hx[2 * r] != hx[2 * r] || hx[2 * r + 1] != hx[2 * r + 1]) { wmax = nan(""); break; }
    wmin = fmin(wmin, hx[2 * r]); wmax = fmax(wmax, hx[2 * r + 1]);
  }
  if (wmin < 0.0) return fail(MCM_ERR_INVALID, "mcm_resample_systematic_sharded: weights must be non-negative");
  if (!(wmax > 0.0) || !isfinite(wmax)) return fail(MCM_ERR_INVALID, "mcm_resample_systematic_sharded: weights must be finite with a positive maximum");
  int e;
  frexp(wmax, &e);
  int lg = 0;
  while ((1ull << lg) < n_global) ++lg;
  const double scale = ldexp(1.0, 62 - lg - e);
  // 2. local fixed-point scan and this rank's total
  unsigned long long* W = nullptr;
  unsigned long long S_local = 0;
  const int grid = stream_grid(n_local ? n_local : 1, 2, 8);
  if (n_local) {
    MCM_CUDA(cudaMallocAsync((void**)&W, n_local * sizeof(unsigned long long), c.stream));
    k_weights_to_fixed<<<grid, kThreads, 0, c.stream>>>(w_local, W, n_local, scale, c.d_flags);
    MCM_CHECK_LAUNCH();
    rc = device_scan<unsigned long long, true>(W, n_local, nullptr);
    if (rc) return rc;
    MCM_CUDA(cudaMemcpyAsync(d_x, W + (n_local - 1), 8, cudaMemcpyDeviceToDevice, c.stream));
  } else {
    MCM_CUDA(cudaMemsetAsync(d_x, 0, 8, c.stream));
  }
  // 3. totals of all ranks -> offset of this rank and the global total
  rc = comm_allgather_bytes(d_x, d_x + 2, 8);
  if (rc) return rc;
  std::vector<unsigned long long> tot(world);
  MCM_CUDA(cudaMemcpyAsync(tot.data(), d_x + 2, 8 * (size_t)world, cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d_x, c.stream));
  unsigned long long C0 = 0, S = 0;
  for (int r = 0; r < world; ++r) { if (r < rank) C0 += tot[r]; S += tot[r]; }
  S_local = tot[rank];
  (void)S_local;
  // 4. every source writes its run of outputs to the owners of those positions
  if (n_local) {
    const unsigned long long U = (unsigned long long)floor(u * 4294967296.0);
    const unsigned big_cap = (unsigned)(3 * (n_out / RS_CHUNK) + 64);
    RsBig* d_big = nullptr;
    MCM_CUDA(cudaMallocAsync((void**)&d_big, (size_t)big_cap * sizeof(RsBig) + 16, c.stream));
    unsigned* d_nbig = reinterpret_cast<unsigned*>(d_big + big_cap);
    MCM_CUDA(cudaMemsetAsync(d_nbig, 0, sizeof(unsigned), c.stream));
    RsOutPeer out{};
    for (int r = 0; r < world; ++r) out.win[r] = static_cast<int64_t*>(peers[r]);
    out.per = per_out;
    k_resample_expand<RsOutPeer><<<grid, kThreads, 0, c.stream>>>(W, n_local, U, out, d_big, d_nbig, big_cap, C0, S, n_out, global_lo);
    MCM_CHECK_LAUNCH();
    k_resample_fill<RsOutPeer><<<c.sm_count * 4, kThreads, 0, c.stream>>>(d_big, d_nbig, out, global_lo);
    MCM_CHECK_LAUNCH();
    MCM_CUDA(cudaFreeAsync(d_big, c.stream));
    MCM_CUDA(cudaFreeAsync(W, c.stream));
  }
  return MCM_OK;
}

// ---- any univariate distribution ---------------------------------------------------------------------------------
static int dist_check(const char* who, int dtype, int kind, const void* out, size_t n) {
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "%s: bad dtype %d", who, dtype);
  if (kind < 0 || kind >= mcm::dist::KIND_COUNT) return fail(MCM_ERR_INVALID, "%s: unknown distribution %d", who, kind);
  if (n && !out) return fail(MCM_ERR_INVALID, "%s: null output", who);
  return MCM_OK;
}
template <int MODE>
static int dist_launch(int dtype, int kind, double a, double b, double c3, void* out, size_t n, const int64_t* perm, const void* pin,
                       uint64_t i0, uint64_t n_total, uint64_t seed, uint64_t offset) {
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  const mcm::dist::Params dp{a, b, c3};
  const int grid = stream_grid(n, 1, 8);
  if (dtype == MCM_F64)
    k_dist<double, MODE><<<grid, kThreads, 0, c.stream>>>((double*)out, n, kind, dp, perm, (const double*)pin, i0, n_total, seed, offset, c.d_flags);
  else
    k_dist<float, MODE><<<grid, kThreads, 0, c.stream>>>((float*)out, n, kind, dp, perm, (const float*)pin, i0, n_total, seed, offset, c.d_flags);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
extern "C" int mcm_systematic_dist(int dtype, int kind, double a, double b, double c3, void* out, size_t n,
                                   const int64_t* perm_dev, uint64_t i0, uint64_t n_total) {
  MCM_NEED_INIT();
  if (int rc = dist_check("mcm_systematic_dist", dtype, kind, out, n)) return rc;
  if (n_total == 0 || (!perm_dev && i0 + n > n_total)) return fail(MCM_ERR_INVALID, "mcm_systematic_dist: grid range out of bounds");
  return dist_launch<0>(dtype, kind, a, b, c3, out, n, perm_dev, nullptr, i0, n_total, 0, 0);
}
extern "C" int mcm_rand_dist(int dtype, int kind, double a, double b, double c3, void* out, size_t n, uint64_t seed, uint64_t offset) {
  MCM_NEED_INIT();
  if (int rc = dist_check("mcm_rand_dist", dtype, kind, out, n)) return rc;
  return dist_launch<1>(dtype, kind, a, b, c3, out, n, nullptr, nullptr, 0, 1, seed, offset);
}
extern "C" int mcm_quantile_dist(int dtype, int kind, double a, double b, double c3, const void* p_in, void* out, size_t n) {
  MCM_NEED_INIT();
  if (int rc = dist_check("mcm_quantile_dist", dtype, kind, out, n)) return rc;
  if (n && !p_in) return fail(MCM_ERR_INVALID, "mcm_quantile_dist: null input");
  return dist_launch<2>(dtype, kind, a, b, c3, out, n, nullptr, p_in, 0, 1, 0, 0);
}
