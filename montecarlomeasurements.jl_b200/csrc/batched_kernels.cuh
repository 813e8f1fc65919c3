// batched_kernels.cuh — kernels and dispatch of the per-particle small-matrix layer, shared by batched*.cu
// The reference maps a host closure over the particle index (ℝⁿ2ℝⁿ_function / ℝⁿ2ℂⁿ_function,
// src/particles.jl:241-282; eigvals :395-406; exp / exp! / lyap :603-607; `\` on matrices via bymap): N calls into
// LAPACK on n x n matrices gathered one scalar per particle-variable.  Here entry (r, c) of the uncertain matrix is a
// length-N device vector, thread i owns particle i, loads its n x n matrix with n^2 coalesced reads (adjacent
// threads = adjacent particles of the same entry), runs the algorithm of mcm_small_linalg.cuh in registers / local
// memory and writes every output entry coalesced.  Bound: HBM for n <= 3 (8 (n^2 + outputs) bytes per particle),
// FP64 pipe beyond.
#pragma once
#include "mcm_common.cuh"
#include "mcm_small_linalg.cuh"
#include "mcm_small_linalg_c.cuh"

namespace mcm {

// Two table sizes: up to 8 x 8 the pointer tables fit the classic 4 KiB of kernel parameters; 9..16 use the large
// parameter space of sm_70+ (6.3 KiB here).  Above 8 the loops of mcm_small_linalg.cuh stay rolled and the matrices
// live in local memory (interleaved per thread by the hardware, so still coalesced).
constexpr int LA_MAXN = 16;
constexpr int LA_MAXPTR = LA_MAXN * LA_MAXN;
template <int P>
struct LaArgsT {
  const void* A[P * P];
  const void* B[P * P];
  void* out[P * P];
  void* out_im[P];
};
using LaArgs = LaArgsT<LA_MAXN>;
template <int N> using LaArgsFor = LaArgsT<(N <= 8 ? 8 : LA_MAXN)>;
template <int P>
static LaArgsT<P> la_narrow(const LaArgs& g) {
  LaArgsT<P> r{};
  for (int e = 0; e < P * P; ++e) { r.A[e] = g.A[e]; r.B[e] = g.B[e]; r.out[e] = g.out[e]; }
  for (int e = 0; e < P; ++e) r.out_im[e] = g.out_im[e];
  return r;
}
#define MCM_LA_UNROLL _Pragma("unroll (N <= 8 ? 64 : 1)")

template <typename T> __device__ __forceinline__ double la_ld(const void* p, size_t i) { return (double)static_cast<const T*>(p)[i]; }
template <typename T> __device__ __forceinline__ void la_st(void* p, size_t i, double v) { static_cast<T*>(p)[i] = (T)v; }

template <int OP, int N, typename T>
__global__ void __launch_bounds__(128)
k_batched_la(const __grid_constant__ LaArgsFor<N> g, int m, size_t n_particles, uint32_t* flags) {
  using namespace sla;
  const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n_particles) return;
  Mat<N> a;
MCM_LA_UNROLL
  for (int e = 0; e < N * N; ++e) a.a[e] = la_ld<T>(g.A[e], i);
  bool bad = false;
  if constexpr (OP == MCM_LA_SOLVE) {
    int piv[N];
    bad = lu_factor(a, piv) == 0;
    for (int j = 0; j < m; ++j) {
      double col[N];
MCM_LA_UNROLL
      for (int r = 0; r < N; ++r) col[r] = la_ld<T>(g.B[r * m + j], i);
      lu_solve(a, piv, col, 1);
MCM_LA_UNROLL
      for (int r = 0; r < N; ++r) la_st<T>(g.out[r * m + j], i, col[r]);
    }
  } else if constexpr (OP == MCM_LA_INV) {
    int piv[N];
    bad = lu_factor(a, piv) == 0;
    Mat<N> x;
    set_identity(x, 1.0);
    for (int j = 0; j < N; ++j) lu_solve(a, piv, &x.a[j], N);
MCM_LA_UNROLL
    for (int e = 0; e < N * N; ++e) la_st<T>(g.out[e], i, x.a[e]);
  } else if constexpr (OP == MCM_LA_DET) {
    int piv[N];
    const int s = lu_factor(a, piv);
    la_st<T>(g.out[0], i, s == 0 ? 0.0 : det_from_lu(a, s));
  } else if constexpr (OP == MCM_LA_EXP) {
    bad = !expm(a);
MCM_LA_UNROLL
    for (int e = 0; e < N * N; ++e) la_st<T>(g.out[e], i, a.a[e]);
  } else if constexpr (OP == MCM_LA_EIGVALS) {
    double wr[N], wi[N];
    bad = !eig_general(a, wr, wi);
MCM_LA_UNROLL
    for (int e = 0; e < N; ++e) { la_st<T>(g.out[e], i, wr[e]); la_st<T>(g.out_im[e], i, wi[e]); }
  } else if constexpr (OP == MCM_LA_EIGVALS_SYM) {
    double w[N];
    eig_sym(a, w);
MCM_LA_UNROLL
    for (int e = 0; e < N; ++e) la_st<T>(g.out[e], i, w[e]);
  } else if constexpr (OP == MCM_LA_LYAP) {
    Mat<N> c, x;
MCM_LA_UNROLL
    for (int e = 0; e < N * N; ++e) c.a[e] = la_ld<T>(g.B[e], i);
    bad = !lyap(a, c, x);
MCM_LA_UNROLL
    for (int e = 0; e < N * N; ++e) la_st<T>(g.out[e], i, x.a[e]);
  } else if constexpr (OP == MCM_LA_SVDVALS) {
    double sv[N];
    svdvals(a, sv);
MCM_LA_UNROLL
    for (int e = 0; e < N; ++e) la_st<T>(g.out[e], i, sv[e]);
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
}

template <int OP, int N, typename T>
static int la_launch(const LaArgs& g, int m, size_t n) {
  Ctx& c = ctx();
  const size_t grid = (n + 127) / 128;
  if (grid > 0x7fffffffull) return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg: too many particles");
  k_batched_la<OP, N, T><<<(unsigned)grid, 128, 0, c.stream>>>(la_narrow<(N <= 8 ? 8 : LA_MAXN)>(g), m, n, c.d_flags);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
// sizes LO..HI of one operation (each translation unit instantiates one range: see batched*.cu)
template <int OP, typename T, int N, int HI>
static int la_by_n(int nmat, const LaArgs& g, int m, size_t n) {
  if constexpr (N > HI || (OP == MCM_LA_LYAP && N > 8)) {      // lyap: Kronecker system of order n^2 <= 64 per thread
    return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg: matrix size %d not supported for this operation", nmat);
  } else {
    if (nmat == N) return la_launch<OP, N, T>(g, m, n);
    return la_by_n<OP, T, N + 1, HI>(nmat, g, m, n);
  }
}
template <typename T, int LO, int HI>
static int la_by_op(int op, int nmat, const LaArgs& g, int m, size_t n) {
  switch (op) {
    case MCM_LA_SOLVE: return la_by_n<MCM_LA_SOLVE, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_INV: return la_by_n<MCM_LA_INV, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_DET: return la_by_n<MCM_LA_DET, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_EXP: return la_by_n<MCM_LA_EXP, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_EIGVALS: return la_by_n<MCM_LA_EIGVALS, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_EIGVALS_SYM: return la_by_n<MCM_LA_EIGVALS_SYM, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_LYAP: return la_by_n<MCM_LA_LYAP, T, LO, HI>(nmat, g, m, n);
    case MCM_LA_SVDVALS: return la_by_n<MCM_LA_SVDVALS, T, LO, HI>(nmat, g, m, n);
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg: unknown operation %d", op);
  }
}


// ---- matrices with Complex{Particles} entries, and log of a real uncertain matrix ---------------------------------
// complex.jl:104-181 (ℂⁿ2ℂⁿ_function: exp, exp!, log; ℂⁿ2ℂ_function: det; eigvals; svdvals) and particles.jl:603
// (Base.log(::Matrix{<:AbstractParticles}) = ℝⁿ2ℂⁿ_function(log, p)): entry (r, c) is a pair of length-N vectors.
template <int P>
struct LacArgsT {
  const void* Are[P * P];
  const void* Aim[P * P];     // all null: a real matrix
  void* ore[P * P];
  void* oim[P * P];
};
using LacArgs = LacArgsT<LA_MAXN>;
template <int P>
static LacArgsT<P> lac_narrow(const LacArgs& g) {
  LacArgsT<P> r{};
  for (int e = 0; e < P * P; ++e) { r.Are[e] = g.Are[e]; r.Aim[e] = g.Aim[e]; r.ore[e] = g.ore[e]; r.oim[e] = g.oim[e]; }
  return r;
}
template <int OP, int N, typename T>
__global__ void __launch_bounds__(128)
k_batched_lac(const __grid_constant__ LacArgsT<(N <= 8 ? 8 : LA_MAXN)> g, int real_input, size_t n_particles, uint32_t* flags) {
  using namespace sla;
  const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n_particles) return;
  CMat<N> a;
  for (int e = 0; e < N * N; ++e) a.a[e] = cmk(la_ld<T>(g.Are[e], i), real_input ? 0.0 : la_ld<T>(g.Aim[e], i));
  bool bad = false;
  if constexpr (OP == MCM_LAC_EXP || OP == MCM_LAC_LOG) {
    if constexpr (OP == MCM_LAC_EXP) bad = !cexpm(a); else bad = !clogm(a);
    for (int e = 0; e < N * N; ++e) { la_st<T>(g.ore[e], i, a.a[e].re); la_st<T>(g.oim[e], i, a.a[e].im); }
  } else if constexpr (OP == MCM_LAC_DET) {
    int piv[N];
    const int s = clu_factor(a, piv);
    const cx d = s == 0 ? cmk(0.0, 0.0) : cdet_from_lu(a, s);
    la_st<T>(g.ore[0], i, d.re); la_st<T>(g.oim[0], i, d.im);
  } else if constexpr (OP == MCM_LAC_EIGVALS) {
    cx w[N];
    bad = !ceigvals(a, w);
    for (int e = 0; e < N; ++e) { la_st<T>(g.ore[e], i, w[e].re); la_st<T>(g.oim[e], i, w[e].im); }
  } else if constexpr (OP == MCM_LAC_SVDVALS) {
    double sv[N];
    csvdvals(a, sv);
    for (int e = 0; e < N; ++e) la_st<T>(g.ore[e], i, sv[e]);
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
}
template <int OP, int N, typename T>
static int lac_launch(const LacArgs& g, int real_input, size_t n) {
  Ctx& c = ctx();
  const size_t grid = (n + 127) / 128;
  if (grid > 0x7fffffffull) return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg_complex: too many particles");
  k_batched_lac<OP, N, T><<<(unsigned)grid, 128, 0, c.stream>>>(lac_narrow<(N <= 8 ? 8 : LA_MAXN)>(g), real_input, n, c.d_flags);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
template <int OP, typename T, int N, int HI>
static int lac_by_n(int nmat, const LacArgs& g, int real_input, size_t n) {
  if constexpr (N > HI) {
    return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg_complex: matrix size %d not supported", nmat);
  } else {
    if (nmat == N) return lac_launch<OP, N, T>(g, real_input, n);
    return lac_by_n<OP, T, N + 1, HI>(nmat, g, real_input, n);
  }
}
template <typename T, int LO, int HI>
static int lac_by_op(int op, int nmat, const LacArgs& g, int real_input, size_t n) {
  switch (op) {
    case MCM_LAC_EXP: return lac_by_n<MCM_LAC_EXP, T, LO, HI>(nmat, g, real_input, n);
    case MCM_LAC_LOG: return lac_by_n<MCM_LAC_LOG, T, LO, HI>(nmat, g, real_input, n);
    case MCM_LAC_DET: return lac_by_n<MCM_LAC_DET, T, LO, HI>(nmat, g, real_input, n);
    case MCM_LAC_EIGVALS: return lac_by_n<MCM_LAC_EIGVALS, T, LO, HI>(nmat, g, real_input, n);
    case MCM_LAC_SVDVALS: return lac_by_n<MCM_LAC_SVDVALS, T, LO, HI>(nmat, g, real_input, n);
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: unknown operation %d", op);
  }
}


// one entry per translation unit (the instantiations take minutes in one file: 392 kernels)
int la_run_small(int dtype, int op, int nmat, const LaArgs& g, int m, size_t n);        // batched.cu: n = 1..8
int la_run_big(int dtype, int op, int nmat, const LaArgs& g, int m, size_t n);          // batched_big.cu: n = 9..16
int lac_run_small(int dtype, int op, int nmat, const LacArgs& g, int real_input, size_t n);   // batched_c.cu: n = 1..8
int lac_run_mid(int dtype, int op, int nmat, const LacArgs& g, int real_input, size_t n);     // batched_c_mid.cu: n = 9..12
int lac_run_big(int dtype, int op, int nmat, const LacArgs& g, int real_input, size_t n);     // batched_c_big.cu: n = 13..16

}  // namespace mcm
