// pgemm.cu — a constant matrix applied to a vector of uncertain variables:
//   y_j = bias_j + sum_k A[j,k] * (x_k - shift_k),   j < Mo, k < Mi, every particle i < n.
// Reference: transform_moments `(m .+ (cholesky(S).L/xl) * X')'` (src/sigmapoints.jl:63-72) and the
// Particle BLAS `mul!(y, A, b::Vector{Particles})` = `Matrix(b) * A'` (src/particles.jl:688-707),
// `A * p` (`_pgemv`, :620-631).  The reference copies the particles into an N x Mi matrix, calls
// BLAS gemm and copies the columns back out; here the Mi inputs and Mo outputs stay where they are
// (separate device arrays, particle index contiguous).
//
// Shape: the contraction runs over the VARIABLES (k <= a few thousand), the free dimension is the
// particle index (1e5 .. 1e8).  CTA tile = 128 outputs x 128 particles, k-slabs of 16:
//   A slab   128 x 16 coefficients from the device copy of A (L2 resident), cp.async;
//   X slab    16 x 128 particles: 16 coalesced 1 KiB rows, cp.async, 4-stage ring;
//   shift_k is subtracted when the B fragment is read from shared memory (the centred matrix the
//   reference forms with `X .- mean(X, dims=1)` never exists);
//   FP64 tensor pipe: mma.sync.m16n8k16.f64 (SASS DMMA), warp tile 32 outputs x 64 particles;
//   epilogue adds bias_j and stores 16-byte pairs (64 B contiguous per output row per 4 lanes).
// j-tiles whose rows of A are zero beyond a column (triangular factors) stop their k loop early.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"

namespace mcm {

constexpr int PG_BM = 128;          // outputs per CTA
constexpr int PG_BN = 128;          // particles per CTA
constexpr int PG_BK = 16;           // variables per slab
constexpr int PG_LDA = PG_BK + 4;   // padded A row (doubles)
constexpr int PG_STAGES = 4;
constexpr int PG_T = 256;

__device__ __forceinline__ void pg_dmma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void pg_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void pg_cp8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void pg_cp4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void pg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void pg_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct PgSmem {
  static constexpr int LDX = PG_BN + (sizeof(T) == 8 ? 4 : 8);  // padded X row: conflict-free fragment reads
  static constexpr size_t A_BYTES = (size_t)PG_BM * PG_LDA * sizeof(double);  // 20 KiB
  static constexpr size_t X_BYTES = (size_t)PG_BK * LDX * sizeof(T);
  static constexpr size_t S_BYTES = (size_t)PG_BK * sizeof(double);
  static constexpr size_t STAGE = A_BYTES + ((X_BYTES + 15) & ~(size_t)15) + S_BYTES;
  static constexpr size_t BYTES = STAGE * PG_STAGES;
};

// A: device, row-major Mo x lda (lda = Mi padded to a multiple of 16, zero filled, rows padded to a
// multiple of 128 with zeros).  kend[jt]: number of leading columns of A that can be non-zero for
// j-tile jt, rounded up to a multiple of 16.
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(PG_T)
k_pgemm(const double* __restrict__ A, int lda, const int* __restrict__ kend, const T* const* __restrict__ xs,
        const double* __restrict__ shift, const double* __restrict__ bias, T* const* __restrict__ outs, int Mo, int Mi,
        size_t n, int njt) {
  using SM = PgSmem<T>;
  constexpr int PG_LDX = SM::LDX;
  extern __shared__ __align__(16) unsigned char smem[];
  const int jt = blockIdx.x % njt;                 // j-tile fastest: consecutive CTAs share the X slab in L2
  const size_t i0 = (size_t)(blockIdx.x / njt) * PG_BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 1, wc = warp & 1;         // warp tile: outputs wr*32.., particles wc*64..
  const int g = lane >> 2, t4 = lane & 3;
  const int nk = kend[jt] / PG_BK;
  const bool full_cols = (i0 + PG_BN <= n);

  double accum[2][8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) accum[i][j][e] = 0.0;

  auto issue = [&](int slab) {
    unsigned char* st = smem + (size_t)(slab % PG_STAGES) * SM::STAGE;
    double* sA = reinterpret_cast<double*>(st);
    T* sX = reinterpret_cast<T*>(st + SM::A_BYTES);
    double* sS = reinterpret_cast<double*>(st + SM::A_BYTES + ((SM::X_BYTES + 15) & ~(size_t)15));
    const int k0 = slab * PG_BK;
    // A slab: 128 rows x 16 doubles = 1024 16-byte pieces, 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = tid + PG_T * q;
      const int row = p >> 3, kk = (p & 7) * 2;
      pg_cp16(sA + row * PG_LDA + kk, A + (size_t)(jt * PG_BM + row) * lda + k0 + kk);
    }
    // X slab: 16 rows x 128 particles
    constexpr int EPV = 16 / (int)sizeof(T);          // elements per 16 bytes
    constexpr int PIECES = PG_BK * PG_BN / EPV;        // 1024 (f64) / 512 (f32)
#pragma unroll
    for (int q = 0; q < PIECES / PG_T; ++q) {
      const int p = tid + PG_T * q;
      const int row = p / (PG_BN / EPV), col = (p % (PG_BN / EPV)) * EPV;
      T* d = sX + row * PG_LDX + col;
      const int k = k0 + row;
      const T* src = k < Mi ? xs[k] : nullptr;
      if (src && full_cols) {
        if constexpr (ALIGNED) pg_cp16(d, src + i0 + col);
        else {
#pragma unroll
          for (int e = 0; e < EPV; ++e) {
            if constexpr (sizeof(T) == 8) pg_cp8(d + e, src + i0 + col + e); else pg_cp4(d + e, src + i0 + col + e);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPV; ++e) d[e] = (src && i0 + col + e < n) ? src[i0 + col + e] : T(0);
      }
    }
    if (tid < PG_BK) sS[tid] = (shift && k0 + tid < Mi) ? shift[k0 + tid] : 0.0;
  };

#pragma unroll 1
  for (int s = 0; s < PG_STAGES - 1; ++s) {
    if (s < nk) issue(s);
    pg_commit();
  }
  for (int s = 0; s < nk; ++s) {
    pg_wait<PG_STAGES - 2>();
    __syncthreads();
    if (s + PG_STAGES - 1 < nk) issue(s + PG_STAGES - 1);
    pg_commit();
    const unsigned char* st = smem + (size_t)(s % PG_STAGES) * SM::STAGE;
    const double* sA = reinterpret_cast<const double*>(st);
    const T* sX = reinterpret_cast<const T*>(st + SM::A_BYTES);
    const double* sS = reinterpret_cast<const double*>(st + SM::A_BYTES + ((SM::X_BYTES + 15) & ~(size_t)15));
    double af[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int v = 0; v < 8; ++v) af[i][v] = sA[(wr * 32 + i * 16 + g + 8 * (v & 1)) * PG_LDA + t4 + 4 * (v >> 1)];
    double sh[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) sh[v] = sS[t4 + 4 * v];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double bf[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) bf[v] = (double)sX[(t4 + 4 * v) * PG_LDX + wc * 64 + j * 8 + g] - sh[v];
#pragma unroll
      for (int i = 0; i < 2; ++i) pg_dmma(accum[i][j], af[i], bf);
    }
  }
  pg_wait<0>();

  // C fragment e of block (i,j): output row wr*32 + i*16 + g + 8*(e>>1), particle wc*64 + j*8 + 2*t4 + (e&1)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = jt * PG_BM + wr * 32 + i * 16 + g + 8 * h;
      if (r >= Mo) continue;
      const double bj = bias ? bias[r] : 0.0;
      T* __restrict__ o = outs[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const size_t col = i0 + wc * 64 + j * 8 + 2 * t4;
        const double v0 = accum[i][j][2 * h] + bj, v1 = accum[i][j][2 * h + 1] + bj;
        if (ALIGNED && col + 1 < n) {
          if constexpr (sizeof(T) == 8) __stcs(reinterpret_cast<double2*>(o + col), make_double2(v0, v1));
          else __stcs(reinterpret_cast<float2*>(o + col), make_float2((float)v0, (float)v1));
        } else {
          if (col < n) o[col] = (T)v0;
          if (col + 1 < n) o[col + 1] = (T)v1;
        }
      }
    }
}

template <typename T>
int pgemm_tensor_impl(const double* A_host, int Mo, int Mi, const T* const* xs_host, const double* shift_host,
                      const double* bias_host, T* const* outs_host, size_t n, const double* A_dev);
static int g_pgemm_use_tensor = 1;
static int g_pgemm_last_engine = 0;   // 0 = FP64 DMMA, 1 = tcgen05 int8 (Ozaki)

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_pgemm_use_tensor(int on) { g_pgemm_use_tensor = on; return MCM_OK; }

namespace mcm {
// The tensor path of mcm_pgemm for a coefficient matrix that already lives on the device (mcm_transform_moments: the
// factor product never visits the host: no 8 MB download, host triangle pass, finiteness scan, staging copy and upload
// at C5).  Returns MCM_OK (done), -1 (not taken: the caller downloads A and calls mcm_pgemm) or an error.
int pgemm_tensor_device_A(int dtype, const double* A_dev, int Mo, int Mi, const void* const* xs, const double* shift_host,
                          const double* bias_host, void* const* outs, size_t n) {
  if (!g_pgemm_use_tensor || Mo < 64 || Mi < 128 || n < 8192) return -1;
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  for (int k = 0; k < Mi; ++k)
    if (!xs[k] || reinterpret_cast<uintptr_t>(xs[k]) % esz) return -1;
  for (int j = 0; j < Mo; ++j) {
    if (!outs[j] || reinterpret_cast<uintptr_t>(outs[j]) % esz) return -1;
    for (int k = 0; k < Mi; ++k)
      if (outs[j] == xs[k]) return -1;
  }
  const int rc = dtype == MCM_F64
                     ? pgemm_tensor_impl<double>(nullptr, Mo, Mi, (const double* const*)xs, shift_host, bias_host, (double* const*)outs, n, A_dev)
                     : pgemm_tensor_impl<float>(nullptr, Mo, Mi, (const float* const*)xs, shift_host, bias_host, (float* const*)outs, n, A_dev);
  if (rc == MCM_OK) g_pgemm_last_engine = 1;
  return rc;
}
}  // namespace mcm
extern "C" int mcm_pgemm_last_engine(void) { return g_pgemm_last_engine; }

extern "C" int mcm_pgemm(int dtype, const double* A_host, int Mo, int Mi, const void* const* xs, const double* shift_host,
                         const double* bias_host, void* const* outs, size_t n) {
  MCM_NEED_INIT();
  if (!A_host || !xs || !outs || Mo < 1 || Mi < 1) return fail(MCM_ERR_INVALID, "mcm_pgemm: bad arguments");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_pgemm: bad dtype %d", dtype);
  if (n == 0) return MCM_OK;
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  bool aligned = true;
  for (int k = 0; k < Mi; ++k) {
    if (!xs[k]) return fail(MCM_ERR_INVALID, "mcm_pgemm: input %d is null", k);
    if (reinterpret_cast<uintptr_t>(xs[k]) % esz) return fail(MCM_ERR_INVALID, "mcm_pgemm: input %d is misaligned", k);
    aligned &= (reinterpret_cast<uintptr_t>(xs[k]) & 15u) == 0;
  }
  for (int j = 0; j < Mo; ++j) {
    if (!outs[j]) return fail(MCM_ERR_INVALID, "mcm_pgemm: output %d is null", j);
    if (reinterpret_cast<uintptr_t>(outs[j]) % esz) return fail(MCM_ERR_INVALID, "mcm_pgemm: output %d is misaligned", j);
    aligned &= (reinterpret_cast<uintptr_t>(outs[j]) & 15u) == 0;
    for (int k = 0; k < Mi; ++k)
      if (outs[j] == xs[k]) return fail(MCM_ERR_INVALID, "mcm_pgemm: outputs must not alias inputs");
  }
  // large problems: tcgen05 / TMEM path (pgemm_tc.cu); it declines inputs it cannot represent safely
  g_pgemm_last_engine = 0;
  if (g_pgemm_use_tensor && Mo >= 64 && Mi >= 128 && n >= 8192) {
    const int rc = dtype == MCM_F64
                       ? pgemm_tensor_impl<double>(A_host, Mo, Mi, (const double* const*)xs, shift_host, bias_host, (double* const*)outs, n, nullptr)
                       : pgemm_tensor_impl<float>(A_host, Mo, Mi, (const float* const*)xs, shift_host, bias_host, (float* const*)outs, n, nullptr);
    if (rc == MCM_OK) { g_pgemm_last_engine = 1; return MCM_OK; }
    if (rc != -1) return rc;
  }
  Ctx& c = ctx();
  const int njt = (Mo + PG_BM - 1) / PG_BM;
  const int lda = (Mi + PG_BK - 1) / PG_BK * PG_BK;
  const size_t nit = (n + PG_BN - 1) / PG_BN;
  if (nit * (size_t)njt > 0x7fffffffull) return fail(MCM_ERR_UNSUPPORTED, "mcm_pgemm: more than 2^31 tiles");
  // host staging: padded A, kend, pointer tables, shift, bias — one upload
  const size_t a_bytes = (size_t)njt * PG_BM * lda * sizeof(double);
  const size_t kend_bytes = ((size_t)njt * sizeof(int) + 15) & ~(size_t)15;
  const size_t tab_bytes = ((size_t)(Mi + Mo) * sizeof(void*) + 15) & ~(size_t)15;
  const size_t sh_bytes = (size_t)Mi * sizeof(double), bi_bytes = (size_t)Mo * sizeof(double);
  std::vector<unsigned char> host(a_bytes + kend_bytes + tab_bytes + sh_bytes + bi_bytes, 0);
  double* hA = reinterpret_cast<double*>(host.data());
  int* hk = reinterpret_cast<int*>(host.data() + a_bytes);
  for (int j = 0; j < Mo; ++j) memcpy(hA + (size_t)j * lda, A_host + (size_t)j * Mi, (size_t)Mi * sizeof(double));
  for (int t = 0; t < njt; ++t) {
    int last = 0;
    for (int j = t * PG_BM; j < std::min(Mo, (t + 1) * PG_BM); ++j)
      for (int k = Mi; k > last; --k)
        if (A_host[(size_t)j * Mi + k - 1] != 0.0) { last = k; break; }
    hk[t] = (last + PG_BK - 1) / PG_BK * PG_BK;
  }
  unsigned char* htab = host.data() + a_bytes + kend_bytes;
  memcpy(htab, xs, (size_t)Mi * sizeof(void*));
  memcpy(htab + (size_t)Mi * sizeof(void*), outs, (size_t)Mo * sizeof(void*));
  if (shift_host) memcpy(host.data() + a_bytes + kend_bytes + tab_bytes, shift_host, sh_bytes);
  if (bias_host) memcpy(host.data() + a_bytes + kend_bytes + tab_bytes + sh_bytes, bias_host, bi_bytes);
  unsigned char* dev = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&dev, host.size(), c.stream));
  MCM_CUDA(cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));  // staging vector dies at return
  const double* dA = reinterpret_cast<const double*>(dev);
  const int* dk = reinterpret_cast<const int*>(dev + a_bytes);
  void** dtab = reinterpret_cast<void**>(dev + a_bytes + kend_bytes);
  const double* dsh = shift_host ? reinterpret_cast<const double*>(dev + a_bytes + kend_bytes + tab_bytes) : nullptr;
  const double* dbi = bias_host ? reinterpret_cast<const double*>(dev + a_bytes + kend_bytes + tab_bytes + sh_bytes) : nullptr;
  static bool attr = false;
  if (!attr) {
    MCM_CUDA(cudaFuncSetAttribute(k_pgemm<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PgSmem<double>::BYTES));
    MCM_CUDA(cudaFuncSetAttribute(k_pgemm<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PgSmem<double>::BYTES));
    MCM_CUDA(cudaFuncSetAttribute(k_pgemm<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PgSmem<float>::BYTES));
    MCM_CUDA(cudaFuncSetAttribute(k_pgemm<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PgSmem<float>::BYTES));
    attr = true;
  }
  const unsigned grid = (unsigned)(nit * (size_t)njt);
  if (dtype == MCM_F64) {
    if (aligned) k_pgemm<double, true><<<grid, PG_T, PgSmem<double>::BYTES, c.stream>>>(dA, lda, dk, (const double* const*)dtab, dsh, dbi, (double* const*)(dtab + Mi), Mo, Mi, n, njt);
    else k_pgemm<double, false><<<grid, PG_T, PgSmem<double>::BYTES, c.stream>>>(dA, lda, dk, (const double* const*)dtab, dsh, dbi, (double* const*)(dtab + Mi), Mo, Mi, n, njt);
  } else {
    if (aligned) k_pgemm<float, true><<<grid, PG_T, PgSmem<float>::BYTES, c.stream>>>(dA, lda, dk, (const float* const*)dtab, dsh, dbi, (float* const*)(dtab + Mi), Mo, Mi, n, njt);
    else k_pgemm<float, false><<<grid, PG_T, PgSmem<float>::BYTES, c.stream>>>(dA, lda, dk, (const float* const*)dtab, dsh, dbi, (float* const*)(dtab + Mi), Mo, Mi, n, njt);
  }
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(dev, c.stream));
  return MCM_OK;
}
