// gram.cu — pcov(v::Vector{Particles}) = cov(Matrix(v))  (src/particles.jl:476, 461-465).
//
// Matrix(v) is N x M column-major: column j is variable j, contiguous over the
// particle index.  Statistics.cov centres the columns, forms X'X and scales by
// 1/(N-1).  Here the M variables stay where they are (M separate device arrays,
// each K = N particles long, K-contiguous) — exactly the "A * A^T with A
// row-major" shape, so both MMA operands are read K-major with coalesced
// 128-bit loads and the centring is fused into the operand staging (no N x M
// copy, no centred copy: the reference makes two).
//
// Float64 path: tcgen05 has no f64 kind, so the contraction runs on the FP64
// tensor pipe through mma.sync.m8n8k4.f64 (DMMA), 128x128x16 CTA tiles,
// upper-triangular tile set (symmetry), split-K with a fixed-order second pass
// when there are too few tiles to fill 148 SMs.
#include "mcm_common.cuh"

namespace mcm {

constexpr int BM = 128;      // CTA tile (rows and cols of C)
constexpr int BK = 16;       // k-slab
constexpr int LDS = BK + 4;  // padded smem row (conflict-free 64-bit fragment loads)
constexpr int GT = 256;      // threads: 8 warps as 4 (rows) x 2 (cols), warp tile 32 x 64

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// column means (Statistics.mean per column): grid (M, nsplit); CTA (j, s) sums a
// contiguous slice of variable j with strided per-thread partials + a fixed tree,
// the partial sums are combined in slice order by k_colmeans_finish.
template <typename T>
__global__ void __launch_bounds__(256)
k_colmeans(const T* const* __restrict__ vars, size_t n, size_t slice, double* __restrict__ partial) {
  const T* x = vars[blockIdx.x];
  const size_t b = (size_t)blockIdx.y * slice;
  size_t e = b + slice;
  if (e > n) e = n;
  double s = 0.0;
  for (size_t i = b + threadIdx.x; i < e; i += 256) s += (double)x[i];
  __shared__ double sh[8];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[(size_t)blockIdx.x * gridDim.y + blockIdx.y] = t;
  }
}
__global__ void k_colmeans_finish(const double* __restrict__ partial, int M, int nsplit, size_t n,
                                  double* __restrict__ means) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  double t = 0.0;
  for (int s = 0; s < nsplit; ++s) t += partial[(size_t)j * nsplit + s];
  means[j] = t / (double)n;
}

template <typename T>
static int colmeans_device(const T* const* d_vars, int M, size_t n, double* d_means) {
  Ctx& c = ctx();
  int nsplit = (4 * c.sm_count + M - 1) / M;
  const size_t max_split = (n + 65535) / 65536;
  if ((size_t)nsplit > max_split) nsplit = (int)max_split;
  if (nsplit < 1) nsplit = 1;
  size_t slice = (n + nsplit - 1) / nsplit;
  nsplit = (int)((n + slice - 1) / slice);
  double* partial = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&partial, (size_t)M * nsplit * sizeof(double), c.stream));
  k_colmeans<T><<<dim3(M, nsplit), 256, 0, c.stream>>>(d_vars, n, slice, partial);
  MCM_CHECK_LAUNCH();
  k_colmeans_finish<<<(M + 127) / 128, 128, 0, c.stream>>>(partial, M, nsplit, n, d_means);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(partial, c.stream));
  return MCM_OK;
}

// tile index -> (bi, bj) with bi <= bj over the upper triangle, row-major
__device__ __forceinline__ void tri_decode(int t, int nb, int& bi, int& bj) {
  int i = 0, rem = t;
  while (rem >= nb - i) { rem -= nb - i; ++i; }
  bi = i; bj = i + rem;
}

// C_partial[tile] (+)= sum_k (x_i[k]-mu_i)(x_j[k]-mu_j) over k in this CTA's K range.
// out: if ksplit == 1 -> scaled result straight into C (both triangles);
//      else raw partial tile into ws[(split*ntiles + tile) * BM*BM].
template <typename T>
__global__ void __launch_bounds__(GT)
k_gram(const T* const* __restrict__ vars, const double* __restrict__ means, int M, size_t n, int nb,
       int ksplit, size_t kchunk, double scale, double* __restrict__ C, double* __restrict__ ws) {
  extern __shared__ __align__(16) double smem_dyn[];
  double(*sA)[BM][LDS] = reinterpret_cast<double(*)[BM][LDS]>(smem_dyn);
  double(*sB)[BM][LDS] = reinterpret_cast<double(*)[BM][LDS]>(smem_dyn + 2 * BM * LDS);
  const int tile = blockIdx.x, split = blockIdx.y;
  int bi, bj;
  tri_decode(tile, nb, bi, bj);
  const bool diag = (bi == bj);
  const size_t k_begin = (size_t)split * kchunk;
  size_t k_end = k_begin + kchunk;
  if (k_end > n) k_end = n;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 1, wc = warp & 1;  // warp tile origin: rows wr*32, cols wc*64
  const int g = lane >> 2, t4 = lane & 3;

  // staging assignment: 128 rows x 16 k per operand = 1024 pairs; 4 pairs/thread/operand
  // pair p = tid + 256*q : row = p / 8, kk = (p % 8) * 2
  double accum[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accum[i][j][0] = accum[i][j][1] = 0.0;

  double ra[4][2], rb[4][2];
  auto load_regs = [&](size_t k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = tid + GT * q;
      const int row = p >> 3, kk = (p & 7) * 2;
      const size_t k = k0 + kk;
      {
        const int gr = bi * BM + row;
        double v0 = 0.0, v1 = 0.0;
        if (gr < M) {
          const T* x = vars[gr];
          const double mu = means[gr];
          if (k + 1 < k_end) {
            if constexpr (sizeof(T) == 8) {
              const double2 t = __ldg(reinterpret_cast<const double2*>(x + k));
              v0 = t.x - mu; v1 = t.y - mu;
            } else {
              const float2 t = __ldg(reinterpret_cast<const float2*>(x + k));
              v0 = (double)t.x - mu; v1 = (double)t.y - mu;
            }
          } else if (k < k_end) {
            v0 = (double)x[k] - mu;
          }
        }
        ra[q][0] = v0; ra[q][1] = v1;
      }
      if (!diag) {
        const int gr = bj * BM + row;
        double v0 = 0.0, v1 = 0.0;
        if (gr < M) {
          const T* x = vars[gr];
          const double mu = means[gr];
          if (k + 1 < k_end) {
            if constexpr (sizeof(T) == 8) {
              const double2 t = __ldg(reinterpret_cast<const double2*>(x + k));
              v0 = t.x - mu; v1 = t.y - mu;
            } else {
              const float2 t = __ldg(reinterpret_cast<const float2*>(x + k));
              v0 = (double)t.x - mu; v1 = (double)t.y - mu;
            }
          } else if (k < k_end) {
            v0 = (double)x[k] - mu;
          }
        }
        rb[q][0] = v0; rb[q][1] = v1;
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = tid + GT * q;
      const int row = p >> 3, kk = (p & 7) * 2;
      sA[buf][row][kk] = ra[q][0];
      sA[buf][row][kk + 1] = ra[q][1];
      if (!diag) {
        sB[buf][row][kk] = rb[q][0];
        sB[buf][row][kk + 1] = rb[q][1];
      }
    }
  };

  const size_t nk = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;
  if (nk > 0) {
    load_regs(k_begin);
    store_smem(0);
  }
  __syncthreads();
  for (size_t s = 0; s < nk; ++s) {
    const int buf = (int)(s & 1);
    if (s + 1 < nk) load_regs(k_begin + (s + 1) * BK);
    const double(*A)[LDS] = sA[buf];
    const double(*B)[LDS] = diag ? sA[buf] : sB[buf];
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = A[wr * 32 + i * 8 + g][k4 + t4];
#pragma unroll
      for (int j = 0; j < 8; ++j) bf[j] = B[wc * 64 + j * 8 + g][k4 + t4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(accum[i][j], af[i], bf[j]);
    }
    if (s + 1 < nk) store_smem(buf ^ 1);
    __syncthreads();
  }

  // epilogue: accumulator (i,j): row = wr*32 + i*8 + g, cols = wc*64 + j*8 + t4*2 + {0,1}
  if (ksplit == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = bi * BM + wr * 32 + i * 8 + g;
          const int c = bj * BM + wc * 64 + j * 8 + t4 * 2 + e;
          if (r < M && c < M) {
            const double v = accum[i][j][e] * scale;
            C[(size_t)r * M + c] = v;
            if (!diag) C[(size_t)c * M + r] = v;
          }
        }
  } else {
    double* w = ws + ((size_t)split * gridDim.x + tile) * (size_t)(BM * BM);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = wr * 32 + i * 8 + g;
        const int c = wc * 64 + j * 8 + t4 * 2;
        *reinterpret_cast<double2*>(w + r * BM + c) = make_double2(accum[i][j][0], accum[i][j][1]);
      }
  }
}

// fixed-order reduction of split-K partial tiles
__global__ void __launch_bounds__(256)
k_gram_reduce(const double* __restrict__ ws, int M, int nb, int ntiles, int ksplit, double scale,
              double* __restrict__ C) {
  const int tile = blockIdx.x;
  int bi, bj;
  tri_decode(tile, nb, bi, bj);
  for (int e = threadIdx.x; e < BM * BM; e += 256) {
    double s = 0.0;
    for (int sp = 0; sp < ksplit; ++sp) s += ws[((size_t)sp * ntiles + tile) * (size_t)(BM * BM) + e];
    const int r = bi * BM + e / BM, c = bj * BM + e % BM;
    if (r < M && c < M) {
      const double v = s * scale;
      C[(size_t)r * M + c] = v;
      if (bi != bj) C[(size_t)c * M + r] = v;
    }
  }
}

int read_back(double* host_dst, const double* dev_src, size_t count);

template <typename T>
static int gram_impl(const void* const* vars_host, int M, size_t n, double scale, const double* center_host,
                     double* cov_out_dev, double* means_out_dev) {
  Ctx& c = ctx();
  const T** d_vars = nullptr;
  double* d_means = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_vars, (size_t)M * sizeof(void*), c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_means, (size_t)M * sizeof(double), c.stream));
  // pageable -> device copy is staged before return, so vars_host may be transient
  MCM_CUDA(cudaMemcpyAsync(d_vars, vars_host, (size_t)M * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
  if (center_host) {
    MCM_CUDA(cudaMemcpyAsync(d_means, center_host, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  } else {
    int rc = colmeans_device<T>(d_vars, M, n, d_means);
    if (rc) return rc;
  }
  if (means_out_dev)
    MCM_CUDA(cudaMemcpyAsync(means_out_dev, d_means, (size_t)M * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));

  const int nb = (M + BM - 1) / BM;
  const int ntiles = nb * (nb + 1) / 2;
  int ksplit = 1;
  const int target = 4 * c.sm_count;
  if (ntiles < target) {
    ksplit = (target + ntiles - 1) / ntiles;
    const size_t max_split = (n + 4095) / 4096;  // keep >= 4096 particles per split
    if ((size_t)ksplit > max_split) ksplit = (int)max_split;
    if (ksplit < 1) ksplit = 1;
  }
  size_t kchunk = (n + ksplit - 1) / ksplit;
  kchunk = (kchunk + BK - 1) / BK * BK;  // keep 16-byte alignment of every split start
  ksplit = (int)((n + kchunk - 1) / kchunk);
  double* ws = nullptr;
  if (ksplit > 1)
    MCM_CUDA(cudaMallocAsync((void**)&ws, (size_t)ksplit * ntiles * BM * BM * sizeof(double), c.stream));
  dim3 grid(ntiles, ksplit);
  constexpr size_t kSmem = 4 * BM * LDS * sizeof(double);  // 80 KiB: A and B, double buffered
  static bool attr_set = false;
  if (!attr_set) {
    MCM_CUDA(cudaFuncSetAttribute(k_gram<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    attr_set = true;
  }
  k_gram<T><<<grid, GT, kSmem, c.stream>>>(d_vars, d_means, M, n, nb, ksplit, kchunk, scale, cov_out_dev, ws);
  MCM_CHECK_LAUNCH();
  if (ksplit > 1) {
    k_gram_reduce<<<ntiles, 256, 0, c.stream>>>(ws, M, nb, ntiles, ksplit, scale, cov_out_dev);
    MCM_CHECK_LAUNCH();
    MCM_CUDA(cudaFreeAsync(ws, c.stream));
  }
  MCM_CUDA(cudaFreeAsync(d_vars, c.stream));
  MCM_CUDA(cudaFreeAsync(d_means, c.stream));
  return MCM_OK;
}

int gram_dispatch(int dtype, const void* const* vars, int M, size_t n, double scale, const double* center_host,
                  double* cov_out_dev, double* means_out_dev) {
  if (dtype == MCM_F64) return gram_impl<double>(vars, M, n, scale, center_host, cov_out_dev, means_out_dev);
  if (dtype == MCM_F32) return gram_impl<float>(vars, M, n, scale, center_host, cov_out_dev, means_out_dev);
  return fail(MCM_ERR_INVALID, "mcm_gram: bad dtype %d", dtype);
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_gram(int dtype, const void* const* vars, int M, size_t n, int scale_mode,
                        const double* center_host, double* cov_out_dev, double* means_out_dev) {
  MCM_NEED_INIT();
  if (!vars || !cov_out_dev || M < 1) return fail(MCM_ERR_INVALID, "mcm_gram: null pointer or M < 1");
  if (n < 1) return fail(MCM_ERR_INVALID, "mcm_gram: no particles");
  for (int i = 0; i < M; ++i)
    if (!vars[i] || (reinterpret_cast<uintptr_t>(vars[i]) & 15u))
      return fail(MCM_ERR_INVALID, "mcm_gram: variable %d is null or not 16-byte aligned", i);
  double scale;
  if (scale_mode == 1) scale = 1.0 / ((double)n - 1.0);  // corrected (n == 1 -> NaN like Base: 0/0)
  else if (scale_mode == 0) scale = 1.0 / (double)n;
  else if (scale_mode == -1) scale = 1.0;
  else return fail(MCM_ERR_INVALID, "mcm_gram: scale must be 1, 0 or -1");
  return gram_dispatch(dtype, vars, M, n, scale, center_host, cov_out_dev, means_out_dev);
}

extern "C" int mcm_colmeans(int dtype, const void* const* vars, int M, size_t n, double* means_out_host) {
  MCM_NEED_INIT();
  if (!vars || !means_out_host || M < 1 || n < 1) return fail(MCM_ERR_INVALID, "mcm_colmeans: bad arguments");
  Ctx& c = ctx();
  if ((size_t)M * sizeof(double) > c.pinned_bytes) return fail(MCM_ERR_INVALID, "mcm_colmeans: M too large");
  void** d_vars = nullptr;
  double* d_means = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_vars, (size_t)M * sizeof(void*), c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_means, (size_t)M * sizeof(double), c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_vars, vars, (size_t)M * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
  int rc;
  if (dtype == MCM_F64) rc = colmeans_device<double>((const double* const*)d_vars, M, n, d_means);
  else if (dtype == MCM_F32) rc = colmeans_device<float>((const float* const*)d_vars, M, n, d_means);
  else return fail(MCM_ERR_INVALID, "mcm_colmeans: bad dtype %d", dtype);
  if (rc) return rc;
  rc = read_back(means_out_host, d_means, (size_t)M);
  MCM_CUDA(cudaFreeAsync(d_vars, c.stream));
  MCM_CUDA(cudaFreeAsync(d_means, c.stream));
  return rc;
}
