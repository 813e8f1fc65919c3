// dense.cu — the M x M ("host-size", M <= a few thousand) dense steps of transform_moments
// (src/sigmapoints.jl:63-72):  xl = cholesky(cov(X)).L,  A = cholesky(S).L / xl,  Y = m + A (X - mean)'.
// The reference calls LAPACK potrf twice and a triangular solve; a 1024^2 factorisation on the host cores
// of a GPU box costs 20-30 ms, more than the 2e12-flop Gram and application kernels around it, so both run
// here as blocked device routines (64-column panels, Float64 FMA):
//   potrf  (right-looking): diagonal block in one CTA -> panel solve L21 L11' = A21 -> trailing update A22 -= L21 L21'
//   trsm   X L = B (right side, lower, no transpose), panels right to left: B_p -= X_right L_right,p ; X_p L_pp = B_p
// Matrices are row-major with leading dimension ld; only the lower triangle of a factor is referenced / written
// (the strict upper triangle of the potrf output is zeroed).
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"

namespace mcm {

int pgemm_tensor_device_A(int dtype, const double* A_dev, int Mo, int Mi, const void* const* xs, const double* shift_host,
                          const double* bias_host, void* const* outs, size_t n);   // pgemm.cu

constexpr int DN_NB = 64;   // panel width
static int g_dense_v2 = 2;   // mcm_dense_use_v2: 2 = as 1 with the panelled diagonal block (k_potrf_diag_v3); 1 = register-tiled diagonal block, right-looking solve, chol(S) on a side stream; 0 = first versions
constexpr int DN_T = 256;

// ---- diagonal block: unblocked Cholesky of an nb x nb block held in shared memory (one CTA) -----------------
__global__ void __launch_bounds__(DN_T)
k_potrf_diag(double* __restrict__ A, int ld, int nb, int* __restrict__ info, int col0) {
  // thread (ty, tx) = (t / 16, t % 16) keeps elements (ty + 16 a, tx + 16 b), a, b < 4, in registers; per column k:
  // pivot -> barrier -> owners scale and publish column k -> barrier -> everyone updates its own 16 elements.
  // (first version: block in shared memory, trailing update indexed through a runtime division: 63 us per block)
  __shared__ double colk[DN_NB];
  __shared__ double s_diag;
  __shared__ int s_bad;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double r[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, j = tx + 16 * b;
      r[a][b] = (i < nb && j <= i) ? A[(size_t)i * ld + j] : 0.0;
    }
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const int kq = k >> 4, kr = k & 15;
    if (ty == kr && tx == kr) {
      double akk = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (a == kq) akk = r[a][a];
      if (!(akk > 0.0)) {    // not positive definite (or NaN): report the 1-based column like LAPACK
        atomicCAS(info, 0, col0 + k + 1);
        s_bad = 1;
      } else {
        s_diag = sqrt(akk);
      }
    }
    __syncthreads();
    if (s_bad) return;
    const double d = s_diag;
    if (tx == kr) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (b == kq) {
            const int i = ty + 16 * a;
            if (i == k) r[a][b] = d;
            else if (i > k) r[a][b] /= d;
            if (i >= k) colk[i] = r[a][b];
          }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 16 * a, j = tx + 16 * b;
        if (j > k && j <= i) r[a][b] -= colk[i] * colk[j];
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, j = tx + 16 * b;
      if (i < nb && j < nb) A[(size_t)i * ld + j] = j <= i ? r[a][b] : 0.0;   // upper part of the block: zeros
    }
}

// third version (default, mcm_dense_use_v2(2)): the same arithmetic in the same order — every element sees its
// updates a_ij -= l_ik l_jk with k ascending, every off-diagonal entry one division by the rounded pivot root — but
// organised in panels of 16 columns so that the CTA meets at 3 barriers per panel instead of 2 per column:
//   (1) the 16 x 16 diagonal block is factored by ONE warp, lane i holding row i in registers, pivots and multipliers
//       exchanged by shuffles (no barrier, no shared-memory round trip inside the 16 columns);
//   (2) the rows below solve against it, one thread per row (forward substitution along the row);
//   (3) all threads apply the panel to the trailing lower triangle.
// ncu showed the second version at 92 us per 64 x 64 block (2750 cycles per column between two CTA-wide barriers).
__global__ void __launch_bounds__(DN_T)
k_potrf_diag_v3(double* __restrict__ A, int ld, int nb, int* __restrict__ info, int col0) {
  __shared__ double s[DN_NB][DN_NB + 1];
  __shared__ int s_bad;
  for (int e = threadIdx.x; e < DN_NB * DN_NB; e += DN_T) {
    const int i = e / DN_NB, j = e % DN_NB;
    s[i][j] = (i < nb && j <= i) ? A[(size_t)i * ld + j] : (i == j ? 1.0 : 0.0);    // identity padding beyond nb
  }
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < nb; c0 += 16) {
    if (warp == 0) {
      const int li = lane & 15;                       // lanes 16..31 mirror lanes 0..15 (their results are not stored)
      double a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = j <= li ? s[c0 + li][c0 + j] : 0.0;
      int bad_col = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double akk = __shfl_sync(0xffffffffu, a[k], k);
        if (!(akk > 0.0) && bad_col == 0) bad_col = col0 + c0 + k + 1;     // not positive definite (or NaN), 1-based like LAPACK
        const double d = sqrt(akk);
        if (li == k) a[k] = d;
        else if (li > k) a[k] /= d;
        const double lik = a[k];
#pragma unroll
        for (int j = k + 1; j < 16; ++j) {
          const double ljk = __shfl_sync(0xffffffffu, lik, j);
          if (li >= j) a[j] -= lik * ljk;
        }
      }
      if (bad_col) {
        if (lane == 0) { atomicCAS(info, 0, bad_col); s_bad = 1; }
      } else if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j <= li) s[c0 + li][c0 + j] = a[j];
      }
    }
    __syncthreads();
    if (s_bad) return;
    // (2) rows below the diagonal block: x L11' = a, one thread per row
    const int r = c0 + 16 + (int)threadIdx.x;
    if (r < nb) {
      double x[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double v = s[r][c0 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) v -= x[k] * s[c0 + j][c0 + k];
        x[j] = v / s[c0 + j][c0 + j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) s[r][c0 + j] = x[j];
    }
    __syncthreads();
    // (3) trailing lower triangle: a_ij -= l_ik l_jk, k ascending
    const int t0 = c0 + 16, m = nb - t0;
    if (m > 0) {
      for (int e = threadIdx.x; e < m * m; e += DN_T) {
        const int i = t0 + e / m, j = t0 + e % m;
        if (j <= i) {
          double v = s[i][j];
#pragma unroll
          for (int k = 0; k < 16; ++k) v -= s[i][c0 + k] * s[j][c0 + k];
          s[i][j] = v;
        }
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < nb * nb; e += DN_T) {
    const int i = e / nb, j = e % nb;
    A[(size_t)i * ld + j] = j <= i ? s[i][j] : 0.0;   // upper part of the block: zeros
  }
}

// first version (kept behind mcm_dense_use_v2(0)): block in shared memory, trailing update through a runtime division
__global__ void __launch_bounds__(DN_T)
k_potrf_diag_v1(double* __restrict__ A, int ld, int nb, int* __restrict__ info, int col0) {
  __shared__ double s[DN_NB][DN_NB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += DN_T) {
    const int i = e / nb, j = e % nb;
    s[i][j] = j <= i ? A[(size_t)i * ld + j] : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const double akk = s[k][k];
    if (!(akk > 0.0)) {      // not positive definite (or NaN): report the 1-based column like LAPACK
      if (threadIdx.x == 0) atomicCAS(info, 0, col0 + k + 1);
      return;
    }
    const double d = sqrt(akk);
    __syncthreads();
    // column k below the diagonal
    for (int i = k + threadIdx.x; i < nb; i += DN_T) s[i][k] = (i == k) ? d : s[i][k] / d;
    __syncthreads();
    // trailing update of the lower triangle
    const int m = nb - k - 1;
    for (int e = threadIdx.x; e < m * m; e += DN_T) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j <= i) s[i][j] -= s[i][k] * s[j][k];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < nb * nb; e += DN_T) {
    const int i = e / nb, j = e % nb;
    A[(size_t)i * ld + j] = s[i][j];   // upper part of the block: zeros
  }
}

// ---- panel solve: rows r of A21 (below the diagonal block): x L11' = a  (forward substitution per row) ---------
__global__ void __launch_bounds__(DN_T)
k_trsm_lt_rows(const double* __restrict__ L11, double* __restrict__ A21, int ld, int nb, int nrows, int g_right_looking) {
  extern __shared__ double dn_smem[];                       // 2 x 64 x 65 doubles (dynamic: above the 48 KiB static limit)
  double (*sl)[DN_NB + 1] = reinterpret_cast<double (*)[DN_NB + 1]>(dn_smem);
  double (*sa)[DN_NB + 1] = reinterpret_cast<double (*)[DN_NB + 1]>(dn_smem + DN_NB * (DN_NB + 1));
  const int r0 = blockIdx.x * DN_NB;
  const int rn = min(DN_NB, nrows - r0);
  for (int e = threadIdx.x; e < nb * nb; e += DN_T) sl[e / nb][e % nb] = L11[(size_t)(e / nb) * ld + e % nb];
  for (int e = threadIdx.x; e < rn * nb; e += DN_T) sa[e / nb][e % nb] = A21[(size_t)(r0 + e / nb) * ld + e % nb];
  __syncthreads();
  if (threadIdx.x < rn) {
    const int r = threadIdx.x;
    if (nb == DN_NB && g_right_looking) {
      // right-looking with the row in registers: once x_k is known it is subtracted from every later entry — 63 - k
      // INDEPENDENT updates instead of the dependent chain of the left-looking form below (41 us per launch, one
      // thread per row).  Every entry still sees its subtractions in ascending k: same bits.
      double a[DN_NB];
#pragma unroll
      for (int j = 0; j < DN_NB; ++j) a[j] = sa[r][j];
#pragma unroll
      for (int k = 0; k < DN_NB; ++k) {
        a[k] = a[k] / sl[k][k];
#pragma unroll
        for (int j = k + 1; j < DN_NB; ++j) a[j] -= a[k] * sl[j][k];
      }
#pragma unroll
      for (int j = 0; j < DN_NB; ++j) sa[r][j] = a[j];
    } else {
      for (int j = 0; j < nb; ++j) {
        double v = sa[r][j];
        for (int k = 0; k < j; ++k) v -= sa[r][k] * sl[j][k];
        sa[r][j] = v / sl[j][j];
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < rn * nb; e += DN_T) A21[(size_t)(r0 + e / nb) * ld + e % nb] = sa[e / nb][e % nb];
}

// ---- C[i,j] -= sum_k A[i,k] * (TRANSB ? B[j,k] : B[k,j]) : 64 x 64 tiles, 4 x 4 outputs per thread -----------------
// LOWER: only tiles with tile-row >= tile-col are computed (symmetric trailing update).
template <bool TRANSB, bool LOWER>
__global__ void __launch_bounds__(DN_T)
k_dgemm_sub(double* __restrict__ Cm, int ldc, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
            int M, int N, int K) {
  if (LOWER && blockIdx.y < blockIdx.x) return;
  __shared__ double sa[16][64 + 1];   // [k][i]
  __shared__ double sb[16][64 + 1];   // [k][j]
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int ti = (threadIdx.x / 16) * 4, tj = (threadIdx.x % 16) * 4;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += DN_T) {
      const int r = e / 16, k = e % 16;                       // A tile: rows i0+r, cols k0+k (row-major: k contiguous)
      sa[k][r] = (i0 + r < M && k0 + k < K) ? A[(size_t)(i0 + r) * lda + k0 + k] : 0.0;
      if (TRANSB) sb[k][r] = (j0 + r < N && k0 + k < K) ? B[(size_t)(j0 + r) * ldb + k0 + k] : 0.0;
    }
    if (!TRANSB) {
      for (int e = threadIdx.x; e < 16 * 64; e += DN_T) {
        const int k = e / 64, c = e % 64;                     // B tile: rows k0+k, cols j0+c (row-major: c contiguous)
        sb[k][c] = (k0 + k < K && j0 + c < N) ? B[(size_t)(k0 + k) * ldb + j0 + c] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = sa[k][ti + u]; b[u] = sb[k][tj + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ti + u, j = j0 + tj + v;
      if (i < M && j < N) Cm[(size_t)i * ldc + j] -= acc[u][v];
    }
}

// ---- row solve x L11 = r (L11 lower, nb x nb): back substitution from the right, 64 rows per CTA ------------------
__global__ void __launch_bounds__(DN_T)
k_trsm_right_rows(const double* __restrict__ L11, double* __restrict__ Bp, int ldl, int ldb, int nb, int nrows) {
  extern __shared__ double dn_smem[];
  double (*sl)[DN_NB + 1] = reinterpret_cast<double (*)[DN_NB + 1]>(dn_smem);
  double (*sx)[DN_NB + 1] = reinterpret_cast<double (*)[DN_NB + 1]>(dn_smem + DN_NB * (DN_NB + 1));
  const int r0 = blockIdx.x * DN_NB;
  const int rn = min(DN_NB, nrows - r0);
  for (int e = threadIdx.x; e < nb * nb; e += DN_T) sl[e / nb][e % nb] = L11[(size_t)(e / nb) * ldl + e % nb];
  for (int e = threadIdx.x; e < rn * nb; e += DN_T) sx[e / nb][e % nb] = Bp[(size_t)(r0 + e / nb) * ldb + e % nb];
  __syncthreads();
  if (threadIdx.x < rn) {
    const int r = threadIdx.x;
    for (int j = nb - 1; j >= 0; --j) {
      double v = sx[r][j];
      for (int k = j + 1; k < nb; ++k) v -= sx[r][k] * sl[k][j];
      sx[r][j] = v / sl[j][j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < rn * nb; e += DN_T) Bp[(size_t)(r0 + e / nb) * ldb + e % nb] = sx[e / nb][e % nb];
}

constexpr size_t DN_TRSM_SMEM = 2 * (size_t)DN_NB * (DN_NB + 1) * sizeof(double);
static int dn_set_attrs() {
  static bool done = false;
  if (!done) {
    MCM_CUDA(cudaFuncSetAttribute(k_trsm_lt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DN_TRSM_SMEM));
    MCM_CUDA(cudaFuncSetAttribute(k_trsm_right_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DN_TRSM_SMEM));
    done = true;
  }
  return MCM_OK;
}

__global__ void k_zero_upper(double* __restrict__ A, int ld, int M) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)M * M) return;
  const int i = (int)(e / M), j = (int)(e % M);
  if (j > i) A[(size_t)i * ld + j] = 0.0;
}
__global__ void k_diag_sqrt(const double* __restrict__ Cm, int M, double* __restrict__ D) {   // D = Diagonal(sqrt.(diag(C)))
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)M * M) return;
  const int i = (int)(e / M), j = (int)(e % M);
  D[e] = i == j ? sqrt(Cm[e]) : 0.0;
}

// in-place lower Cholesky on the library stream; *info_host = 0 or the 1-based column of the first non-positive pivot
// enqueue only (no host synchronisation): *d_info_out receives the device flag to read once the stream has drained
static int potrf_lower_enqueue(double* A, int M, int** d_info_out) {
  Ctx& c = ctx();
  if (int rc = dn_set_attrs()) return rc;
  int* d_info = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_info, sizeof(int), c.stream));
  MCM_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), c.stream));
  *d_info_out = d_info;
  for (int j0 = 0; j0 < M; j0 += DN_NB) {
    const int nb = std::min(DN_NB, M - j0);
    double* A11 = A + (size_t)j0 * M + j0;
    if (g_dense_v2 == 2) k_potrf_diag_v3<<<1, DN_T, 0, c.stream>>>(A11, M, nb, d_info, j0);
    else if (g_dense_v2) k_potrf_diag<<<1, DN_T, 0, c.stream>>>(A11, M, nb, d_info, j0);
    else k_potrf_diag_v1<<<1, DN_T, 0, c.stream>>>(A11, M, nb, d_info, j0);
    MCM_CHECK_LAUNCH();
    const int rest = M - j0 - nb;
    if (rest > 0) {
      double* A21 = A + (size_t)(j0 + nb) * M + j0;
      double* A22 = A + (size_t)(j0 + nb) * M + j0 + nb;
      k_trsm_lt_rows<<<(rest + DN_NB - 1) / DN_NB, DN_T, DN_TRSM_SMEM, c.stream>>>(A11, A21, M, nb, rest, g_dense_v2 == 2 ? 1 : 0);
      MCM_CHECK_LAUNCH();
      const int nt = (rest + 63) / 64;
      k_dgemm_sub<true, true><<<dim3(nt, nt), DN_T, 0, c.stream>>>(A22, M, A21, M, A21, M, rest, rest, nb);
      MCM_CHECK_LAUNCH();
    }
  }
  k_zero_upper<<<(unsigned)(((size_t)M * M + 255) / 256), 256, 0, c.stream>>>(A, M, M);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
static int potrf_lower_finish(int* d_info, int* info_host) {
  Ctx& c = ctx();
  MCM_CUDA(cudaMemcpyAsync(info_host, d_info, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d_info, c.stream));
  return MCM_OK;
}
int potrf_lower_device(double* A, int M, int* info_host) {
  int* d_info = nullptr;
  if (int rc = potrf_lower_enqueue(A, M, &d_info)) return rc;
  return potrf_lower_finish(d_info, info_host);
}

// X L = B, in place in B (nrows x M), L lower triangular M x M
int trsm_right_lower_device(const double* L, double* B, int M, int nrows) {
  Ctx& c = ctx();
  if (int rc = dn_set_attrs()) return rc;
  const int np = (M + DN_NB - 1) / DN_NB;
  for (int p = np - 1; p >= 0; --p) {
    const int j0 = p * DN_NB, nb = std::min(DN_NB, M - j0);
    if (!g_dense_v2) {   // first version, left-looking: B[:, panel] -= X[:, right] * L[right, panel] before the panel solve
      const int right = M - j0 - nb;
      if (right > 0) {
        k_dgemm_sub<false, false><<<dim3((nb + 63) / 64, (nrows + 63) / 64), DN_T, 0, c.stream>>>(
            B + j0, M, B + j0 + nb, M, L + (size_t)(j0 + nb) * M + j0, M, nrows, nb, right);
        MCM_CHECK_LAUNCH();
      }
    }
    k_trsm_right_rows<<<(nrows + DN_NB - 1) / DN_NB, DN_T, DN_TRSM_SMEM, c.stream>>>(L + (size_t)j0 * M + j0, B + j0, M, M, nb, nrows);
    MCM_CHECK_LAUNCH();
    // right-looking: the solved panel updates every column to its left at once, B[:, 0:j0] -= X[:, panel] * L[panel, 0:j0]
    // ((j0/64) x (nrows/64) CTAs with K = 64; the left-looking form ran 16 CTAs over K up to 960: 2.8 ms of 3.6)
    if (g_dense_v2 && j0 > 0) {
      k_dgemm_sub<false, false><<<dim3((j0 + 63) / 64, (nrows + 63) / 64), DN_T, 0, c.stream>>>(
          B, M, B + j0, M, L + (size_t)j0 * M, M, nrows, j0, nb);
      MCM_CHECK_LAUNCH();
    }
  }
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

/* A/B switch for the dense M x M steps: 1 (default) = register-tiled diagonal block, right-looking triangular solve and
 * chol(S) overlapped with the Gram on a side stream; 0 = the first versions.  Same results to rounding. */
extern "C" int mcm_dense_use_v2(int on) { g_dense_v2 = on; return MCM_OK; }

extern "C" int mcm_potrf(double* a_dev, int M, int* info_host) {
  MCM_NEED_INIT();
  if (!a_dev || M < 1 || !info_host) return fail(MCM_ERR_INVALID, "mcm_potrf: bad arguments");
  const int rc = potrf_lower_device(a_dev, M, info_host);
  if (rc) return rc;
  if (*info_host) return fail(MCM_ERR_INVALID, "mcm_potrf: matrix is not positive definite (pivot %d)", *info_host);
  return MCM_OK;
}

extern "C" int mcm_trsm_right_lower(const double* l_dev, double* b_dev, int M, int nrows) {
  MCM_NEED_INIT();
  if (!l_dev || !b_dev || M < 1 || nrows < 0) return fail(MCM_ERR_INVALID, "mcm_trsm_right_lower: bad arguments");
  if (nrows == 0) return MCM_OK;
  return trsm_right_lower_device(l_dev, b_dev, M, nrows);
}

// transform_moments(X, m, S; preserve_latin) (sigmapoints.jl:63-72), everything on device:
//   mu = column means, C = cov(X) [mcm_gram], xl = chol(C).L (or Diagonal(std)), A = chol(S).L / xl, outs = m + A (X - mu)
extern "C" int mcm_transform_moments(int dtype, const void* const* xs, int M, size_t n, const double* m_host,
                                     const double* S_host, int preserve_latin, void* const* outs) {
  MCM_NEED_INIT();
  if (!xs || !outs || !m_host || !S_host || M < 1) return fail(MCM_ERR_INVALID, "mcm_transform_moments: bad arguments");
  if (n < 2) return fail(MCM_ERR_INVALID, "mcm_transform_moments: need at least 2 particles");
  Ctx& c = ctx();
  const size_t mm = (size_t)M * M;
  double* d = nullptr;   // [C | LS | means]
  MCM_CUDA(cudaMallocAsync((void**)&d, (2 * mm + (size_t)M) * sizeof(double), c.stream));
  double *dC = d, *dLS = d + mm, *dmu = d + 2 * mm;
  static cudaStream_t side = nullptr;
  static cudaEvent_t ev_fork = nullptr;
  auto release = [&]() {
    if (side) cudaStreamSynchronize(side);          // the side stream may still be factorising inside this block
    cudaFreeAsync(d, c.stream);
  };
  MCM_CUDA(cudaMemcpyAsync(dLS, S_host, mm * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  // chol(S) does not depend on the particles: its ~100 small launches go to a side stream and run under the Gram's
  // HBM-bound prepass instead of after the Gram (the library stream is swapped for the enqueue only)
  if (!side) {
    MCM_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    MCM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  }
  MCM_CUDA(cudaEventRecord(ev_fork, c.stream));
  MCM_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  int* d_info_S = nullptr;
  const bool overlap = g_dense_v2 != 0;
  if (overlap) {
    const cudaStream_t main_stream = c.stream;
    c.stream = side;
    const int rcS = potrf_lower_enqueue(dLS, M, &d_info_S);
    c.stream = main_stream;
    if (rcS) { cudaStreamSynchronize(side); release(); return rcS; }
  }
  StatsHandoff& ho = stats_handoff();
  ho = StatsHandoff{};
  ho.armed = true;
  auto drop_stats = [&]() {
    if (ho.valid) { cudaFreeAsync(ho.d_ex, c.stream); cudaFreeAsync(ho.d_ratio, c.stream); }
    ho = StatsHandoff{};
  };
  int rc = mcm_gram(dtype, xs, M, n, 1, nullptr, dC, dmu);
  ho.armed = false;
  if (rc) { drop_stats(); release(); return rc; }
  int info = 0;
  if (overlap) {
    const cudaStream_t main_stream = c.stream;
    c.stream = side;
    rc = potrf_lower_finish(d_info_S, &info);      // synchronises the side stream: dLS = chol(S).L from here on
    c.stream = main_stream;
  } else {
    rc = potrf_lower_device(dLS, M, &info);
  }
  if (rc) { drop_stats(); release(); return rc; }
  if (info) { drop_stats(); release(); return fail(MCM_ERR_INVALID, "mcm_transform_moments: the target covariance is not positive definite (pivot %d)", info); }
  if (preserve_latin) {
    double* dD = nullptr;
    MCM_CUDA(cudaMallocAsync((void**)&dD, mm * sizeof(double), c.stream));
    k_diag_sqrt<<<(unsigned)((mm + 255) / 256), 256, 0, c.stream>>>(dC, M, dD);
    MCM_CHECK_LAUNCH();
    MCM_CUDA(cudaMemcpyAsync(dC, dD, mm * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    MCM_CUDA(cudaFreeAsync(dD, c.stream));
  } else {
    rc = potrf_lower_device(dC, M, &info);
    if (rc) { drop_stats(); release(); return rc; }
    if (info) { drop_stats(); release(); return fail(MCM_ERR_INVALID, "mcm_transform_moments: the sample covariance is not positive definite (pivot %d): fewer particles than variables, or dependent columns", info); }
  }
  rc = trsm_right_lower_device(dC, dLS, M, M);   // dLS <- chol(S).L / xl
  if (rc) { drop_stats(); release(); return rc; }
  // Both factors are lower triangular, so is A; its upper part is cleared explicitly (it is exactly zero already; this
  // keeps the early exit of the application kernels independent of that).  The tensor application takes A where it
  // is: only the M means visit the host.  (Until the last session of round 2 A was downloaded, cleared and scanned on
  // the host, copied to pinned staging and uploaded again: ~3 ms of an idle GPU at C5.)
  k_zero_upper<<<(unsigned)((mm + 255) / 256), 256, 0, c.stream>>>(dLS, M, M);
  MCM_CHECK_LAUNCH();
  {
    std::vector<double> mu_dev(M);
    MCM_CUDA(cudaMemcpyAsync(mu_dev.data(), dmu, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    MCM_CUDA(cudaStreamSynchronize(c.stream));
    rc = pgemm_tensor_device_A(dtype, dLS, M, M, xs, mu_dev.data(), m_host, outs, n);
    if (rc != -1) { release(); drop_stats(); return rc; }
  }
  // not taken by the tensor path (small problem, heavy-tailed variable, switched off): the host form of mcm_pgemm.
  // A and the means go back through PINNED staging (a pageable 8 MB copy each way cost ~2 ms here)
  std::vector<double> A_fallback, mu(M);
  double* A = static_cast<double*>(pinned_scratch(0, mm * sizeof(double)));
  if (!A) { A_fallback.resize(mm); A = A_fallback.data(); }
  MCM_CUDA(cudaMemcpyAsync(A, dLS, mm * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaMemcpyAsync(mu.data(), dmu, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  release();
  // both factors are lower triangular, so is A: clear the round-off-free upper part explicitly (it is exactly zero
  // already; this keeps the early-exit of the application kernels independent of that)
  for (int i = 0; i < M; ++i)
    for (int j = i + 1; j < M; ++j) A[(size_t)i * M + j] = 0.0;
  rc = mcm_pgemm(dtype, A, M, M, xs, mu.data(), m_host, outs, n);
  drop_stats();
  return rc;
}
