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
// tensor pipe through mma.sync.m16n8k16.f64 (DMMA), 128x128x16 CTA tiles,
// upper-triangular tile set (symmetry), split-K with a fixed-order second pass
// when there are too few tiles to fill 148 SMs.
#include "mcm_common.cuh"

namespace mcm {

constexpr int BM = 128;      // CTA tile (rows and cols of C)
constexpr int BK = 16;       // k-slab
constexpr int LDS = BK + 4;  // padded smem row (conflict-free 64-bit fragment loads)
constexpr int GT = 256;      // threads: 8 warps as 4 (rows) x 2 (cols), warp tile 32 x 64

// FP64 tensor-pipe MMA, 16x8x16 (sm_90+ shape; one instruction = 2048 FMA).  Fragment layout
// (lane = 4*g + t): A[v]: row g + 8*(v&1), k = t + 4*(v>>1);  B[v]: k = t + 4*v, col g;
// C[v]: row g + 8*(v>>1), col 2*t + (v&1).
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
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
//
// Operand staging: RAW particle values stream global -> shared with cp.async (16 B per request, no
// registers), GSTAGES-deep ring, one __syncthreads per 16-particle slab; the column means are
// subtracted when fragments are read from shared memory (4 + 8 means live in registers), so the
// centred matrix never exists anywhere.  A slab that crosses the end of the K range is written
// synchronously with the tail padded by the row mean (centred value 0).
constexpr int GSTAGES = 4;

template <typename T> struct GramSmem {
  static constexpr int LD = BK + 4;                       // elements per padded row (conflict-free for 4- and 8-byte T)
  static constexpr int STAGE_ELEMS = 2 * BM * LD;         // A rows then B rows
  static constexpr size_t BYTES = (size_t)GSTAGES * STAGE_ELEMS * sizeof(T);
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(GT)
k_gram(const T* const* __restrict__ vars, const double* __restrict__ means, int M, size_t n, int nb,
       int ksplit, size_t kchunk, double scale, double* __restrict__ C, double* __restrict__ ws) {
  using SM = GramSmem<T>;
  constexpr int LD = SM::LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
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

  // ---- loader assignment: 128 rows x 16 k per operand = 1024 pairs of elements; 4 pairs/thread/operand.
  // pair p = tid + 256*q : row = p / 8, kk = (p % 8) * 2.  Row pointers are loop invariant.
  const T* rowA[4];
  const T* rowB[4];
  double fillA[4], fillB[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int row = (tid + GT * q) >> 3;
    const int ga = bi * BM + row, gb = bj * BM + row;
    rowA[q] = ga < M ? vars[ga] : nullptr;
    rowB[q] = gb < M ? vars[gb] : nullptr;
    fillA[q] = ga < M ? means[ga] : 0.0;
    fillB[q] = gb < M ? means[gb] : 0.0;
  }
  // ---- means of the rows this thread's fragments touch (rows beyond M: mean 0, data 0)
  double muA[2][2], muB[8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = bi * BM + wr * 32 + i * 16 + g + 8 * h;
      muA[i][h] = r < M ? means[r] : 0.0;
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = bj * BM + wc * 64 + j * 8 + g;
    muB[j] = r < M ? means[r] : 0.0;
  }

  double accum[2][8][4];  // [m16 block][n8 block][fragment]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) accum[i][j][e] = 0.0;

  const size_t nk = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;

  auto issue = [&](size_t slab) {
    T* st = smem + (size_t)(slab % GSTAGES) * SM::STAGE_ELEMS;
    const size_t k0 = k_begin + slab * BK;
    const bool full = (k0 + BK <= k_end);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = tid + GT * q;
      const int row = p >> 3, kk = (p & 7) * 2;
      T* dA = st + row * LD + kk;
      T* dB = st + (BM + row) * LD + kk;
      if (full) {
        if (rowA[q]) { if constexpr (sizeof(T) == 8) cp_async16(dA, rowA[q] + k0 + kk); else cp_async8(dA, rowA[q] + k0 + kk); }
        else { dA[0] = T(0); dA[1] = T(0); }
        if (!diag) {
          if (rowB[q]) { if constexpr (sizeof(T) == 8) cp_async16(dB, rowB[q] + k0 + kk); else cp_async8(dB, rowB[q] + k0 + kk); }
          else { dB[0] = T(0); dB[1] = T(0); }
        }
      } else {  // tail slab: synchronous, padded with the row mean so the centred value is exactly 0
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const size_t k = k0 + kk + e;
          dA[e] = (rowA[q] && k < k_end) ? rowA[q][k] : (T)fillA[q];
          if (!diag) dB[e] = (rowB[q] && k < k_end) ? rowB[q][k] : (T)fillB[q];
        }
      }
    }
  };
  // NOTE (Float32 tail): (T)fill rounds the mean to Float32, so a padded element contributes
  // (fl32(mu) - mu)^2 ~ 1e-15 * mu^2 — far below the Float32 tolerance.

#pragma unroll 1
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if ((size_t)s < nk) issue(s);
    cp_async_commit();
  }
  for (size_t s = 0; s < nk; ++s) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();  // slab s landed for everyone; everyone finished reading slab s-1
    if (s + GSTAGES - 1 < nk) issue(s + GSTAGES - 1);
    cp_async_commit();
    const T* A = smem + (size_t)(s % GSTAGES) * SM::STAGE_ELEMS;
    const T* B = diag ? A : A + BM * LD;
    double af[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int v = 0; v < 8; ++v)
        af[i][v] = (double)A[(wr * 32 + i * 16 + g + 8 * (v & 1)) * LD + t4 + 4 * (v >> 1)] - muA[i][v & 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double bf[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) bf[v] = (double)B[(wc * 64 + j * 8 + g) * LD + t4 + 4 * v] - muB[j];
#pragma unroll
      for (int i = 0; i < 2; ++i) dmma16816(accum[i][j], af[i], bf);
    }
  }
  cp_async_wait<0>();

  // epilogue: fragment e of block (i,j): row = wr*32 + i*16 + g + 8*(e>>1), col = wc*64 + j*8 + 2*t4 + (e&1)
  if (ksplit == 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = bi * BM + wr * 32 + i * 16 + g + 8 * (e >> 1);
          const int c = bj * BM + wc * 64 + j * 8 + t4 * 2 + (e & 1);
          if (r < M && c < M) {
            const double v = accum[i][j][e] * scale;
            C[(size_t)r * M + c] = v;
            if (!diag) C[(size_t)c * M + r] = v;
          }
        }
  } else {
    double* w = ws + ((size_t)split * gridDim.x + tile) * (size_t)(BM * BM);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = wr * 32 + i * 16 + g + 8 * h;
          const int c = wc * 64 + j * 8 + t4 * 2;
          *reinterpret_cast<double2*>(w + r * BM + c) = make_double2(accum[i][j][2 * h], accum[i][j][2 * h + 1]);
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

// tcgen05 / TMEM path (gram_tc.cu)
template <typename T>
int gram_tensor_impl(const T* const* d_vars, const double* d_means, int M, size_t n, double scale, double* C,
                     const int* pre_ex, const double* pre_ratio);
template <typename T>
int varstats_device(const T* const* d_vars, int M, size_t n, double* d_means, int* d_ex, double* d_ratio);
StatsHandoff& stats_handoff() { static StatsHandoff h; return h; }
static int g_use_tensor = 1;  // mcm_gram_use_tensor: 0 = FP64 DMMA kernel only, 1 = Ozaki int8 on tcgen05 for large problems
static int g_last_engine = 0; // 0 = DMMA, 1 = tcgen05

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
  const bool tensor_ok = g_use_tensor && M >= 256 && n >= 4096;
  int* d_ex = nullptr;
  double* d_ratio = nullptr;
  if (center_host) {
    MCM_CUDA(cudaMemcpyAsync(d_means, center_host, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  } else if (tensor_ok) {
    // one read gives the means AND the row statistics the digit split needs
    MCM_CUDA(cudaMallocAsync((void**)&d_ex, (size_t)M * sizeof(int), c.stream));
    MCM_CUDA(cudaMallocAsync((void**)&d_ratio, (size_t)M * sizeof(double), c.stream));
    int rc = varstats_device<T>(d_vars, M, n, d_means, d_ex, d_ratio);
    if (rc) return rc;
  } else {
    int rc = colmeans_device<T>(d_vars, M, n, d_means);
    if (rc) return rc;
  }
  if (means_out_dev)
    MCM_CUDA(cudaMemcpyAsync(means_out_dev, d_means, (size_t)M * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));

  g_last_engine = 0;
  if (tensor_ok) {
    int rc = gram_tensor_impl<T>(d_vars, d_means, M, n, scale, cov_out_dev, d_ex, d_ratio);
    if (d_ex) {
      StatsHandoff& h = stats_handoff();
      if (h.armed && !h.valid && rc >= 0) {     // mcm_transform_moments: the application reuses them (and frees them)
        h.valid = true; h.M = M; h.n = n; h.first_var = vars_host[0]; h.d_ex = d_ex; h.d_ratio = d_ratio;
      } else {
        MCM_CUDA(cudaFreeAsync(d_ex, c.stream)); MCM_CUDA(cudaFreeAsync(d_ratio, c.stream));
      }
    }
    if (rc >= 0) {  // -1: declined by the accuracy guard (heavy-tailed rows) -> DMMA below
      g_last_engine = 1;
      MCM_CUDA(cudaFreeAsync(d_vars, c.stream));
      MCM_CUDA(cudaFreeAsync(d_means, c.stream));
      return rc;
    }
  }
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
  constexpr size_t kSmem = GramSmem<T>::BYTES;  // 4 stages x (A + B) x 128 rows x 20 elements
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

namespace mcm { extern int g_a_in_tmem; extern int g_oz_last_digits; }
extern "C" int mcm_gram_use_tensor(int on) { g_use_tensor = on; mcm::g_a_in_tmem = (on == 2); return MCM_OK; }
extern "C" int mcm_gram_last_engine(void) { return g_last_engine; }
extern "C" int mcm_gram_last_digits(void) { return mcm::g_oz_last_digits; }

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
