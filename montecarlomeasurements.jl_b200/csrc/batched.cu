// batched.cu — per-particle small-matrix operations on device.
// The reference maps a host closure over the particle index (ℝⁿ2ℝⁿ_function / ℝⁿ2ℂⁿ_function,
// src/particles.jl:241-282; eigvals :395-406; exp / exp! / lyap :603-607; `\` on matrices via bymap): N calls into
// LAPACK on n x n matrices gathered one scalar per particle-variable.  Here entry (r, c) of the uncertain matrix is a
// length-N device vector, thread i owns particle i, loads its n x n matrix with n^2 coalesced reads (adjacent
// threads = adjacent particles of the same entry), runs the algorithm of mcm_small_linalg.cuh in registers / local
// memory and writes every output entry coalesced.  Bound: HBM for n <= 3 (8 (n^2 + outputs) bytes per particle),
// FP64 pipe beyond.
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
template <int OP, typename T>
static int la_by_n(int nmat, const LaArgs& g, int m, size_t n) {
  switch (nmat) {
    case 1: return la_launch<OP, 1, T>(g, m, n);
    case 2: return la_launch<OP, 2, T>(g, m, n);
    case 3: return la_launch<OP, 3, T>(g, m, n);
    case 4: return la_launch<OP, 4, T>(g, m, n);
    default: break;
  }
  if constexpr (OP != MCM_LA_LYAP) {
    switch (nmat) {
      case 5: return la_launch<OP, 5, T>(g, m, n);
      case 6: return la_launch<OP, 6, T>(g, m, n);
      case 7: return la_launch<OP, 7, T>(g, m, n);
      case 8: return la_launch<OP, 8, T>(g, m, n);
      case 9: return la_launch<OP, 9, T>(g, m, n);
      case 10: return la_launch<OP, 10, T>(g, m, n);
      case 11: return la_launch<OP, 11, T>(g, m, n);
      case 12: return la_launch<OP, 12, T>(g, m, n);
      case 13: return la_launch<OP, 13, T>(g, m, n);
      case 14: return la_launch<OP, 14, T>(g, m, n);
      case 15: return la_launch<OP, 15, T>(g, m, n);
      case 16: return la_launch<OP, 16, T>(g, m, n);
      default: break;
    }
  }
  return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg: matrix size %d not supported for this operation", nmat);
}
template <typename T>
static int la_by_op(int op, int nmat, const LaArgs& g, int m, size_t n) {
  switch (op) {
    case MCM_LA_SOLVE: return la_by_n<MCM_LA_SOLVE, T>(nmat, g, m, n);
    case MCM_LA_INV: return la_by_n<MCM_LA_INV, T>(nmat, g, m, n);
    case MCM_LA_DET: return la_by_n<MCM_LA_DET, T>(nmat, g, m, n);
    case MCM_LA_EXP: return la_by_n<MCM_LA_EXP, T>(nmat, g, m, n);
    case MCM_LA_EIGVALS: return la_by_n<MCM_LA_EIGVALS, T>(nmat, g, m, n);
    case MCM_LA_EIGVALS_SYM: return la_by_n<MCM_LA_EIGVALS_SYM, T>(nmat, g, m, n);
    case MCM_LA_LYAP: return la_by_n<MCM_LA_LYAP, T>(nmat, g, m, n);
    case MCM_LA_SVDVALS: return la_by_n<MCM_LA_SVDVALS, T>(nmat, g, m, n);
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
template <int OP, typename T>
static int lac_by_n(int nmat, const LacArgs& g, int real_input, size_t n) {
  switch (nmat) {
    case 1: return lac_launch<OP, 1, T>(g, real_input, n);
    case 2: return lac_launch<OP, 2, T>(g, real_input, n);
    case 3: return lac_launch<OP, 3, T>(g, real_input, n);
    case 4: return lac_launch<OP, 4, T>(g, real_input, n);
    case 5: return lac_launch<OP, 5, T>(g, real_input, n);
    case 6: return lac_launch<OP, 6, T>(g, real_input, n);
    case 7: return lac_launch<OP, 7, T>(g, real_input, n);
    case 8: return lac_launch<OP, 8, T>(g, real_input, n);
    case 9: return lac_launch<OP, 9, T>(g, real_input, n);
    case 10: return lac_launch<OP, 10, T>(g, real_input, n);
    case 11: return lac_launch<OP, 11, T>(g, real_input, n);
    case 12: return lac_launch<OP, 12, T>(g, real_input, n);
    case 13: return lac_launch<OP, 13, T>(g, real_input, n);
    case 14: return lac_launch<OP, 14, T>(g, real_input, n);
    case 15: return lac_launch<OP, 15, T>(g, real_input, n);
    case 16: return lac_launch<OP, 16, T>(g, real_input, n);
    default: return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg_complex: matrix size %d not supported", nmat);
  }
}
template <typename T>
static int lac_by_op(int op, int nmat, const LacArgs& g, int real_input, size_t n) {
  switch (op) {
    case MCM_LAC_EXP: return lac_by_n<MCM_LAC_EXP, T>(nmat, g, real_input, n);
    case MCM_LAC_LOG: return lac_by_n<MCM_LAC_LOG, T>(nmat, g, real_input, n);
    case MCM_LAC_DET: return lac_by_n<MCM_LAC_DET, T>(nmat, g, real_input, n);
    case MCM_LAC_EIGVALS: return lac_by_n<MCM_LAC_EIGVALS, T>(nmat, g, real_input, n);
    case MCM_LAC_SVDVALS: return lac_by_n<MCM_LAC_SVDVALS, T>(nmat, g, real_input, n);
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: unknown operation %d", op);
  }
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_batched_linalg(int op, int dtype, int nmat, int m, const void* const* A, const void* const* B,
                                  void* const* out, void* const* out_im, size_t n) {
  MCM_NEED_INIT();
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: bad dtype %d", dtype);
  if (nmat < 1 || nmat > LA_MAXN) return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg: matrix size must be 1..%d", LA_MAXN);
  if (!A || !out) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: null table");
  const bool needs_b = op == MCM_LA_SOLVE || op == MCM_LA_LYAP;
  if (op == MCM_LA_LYAP) m = nmat;
  if (needs_b && (!B || m < 1 || nmat * m > LA_MAXPTR)) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: bad right-hand side");
  if (op == MCM_LA_EIGVALS && !out_im) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: eigvals needs out_im");
  int nout = 0;
  switch (op) {
    case MCM_LA_SOLVE: nout = nmat * m; break;
    case MCM_LA_INV: case MCM_LA_EXP: case MCM_LA_LYAP: nout = nmat * nmat; break;
    case MCM_LA_DET: nout = 1; break;
    case MCM_LA_EIGVALS: case MCM_LA_EIGVALS_SYM: case MCM_LA_SVDVALS: nout = nmat; break;
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg: unknown operation %d", op);
  }
  LaArgs g{};
  for (int e = 0; e < nmat * nmat; ++e) {
    if (!A[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: A[%d] is null", e);
    g.A[e] = A[e];
  }
  if (needs_b)
    for (int e = 0; e < nmat * m; ++e) {
      if (!B[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: B[%d] is null", e);
      g.B[e] = B[e];
    }
  for (int e = 0; e < nout; ++e) {
    if (!out[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: out[%d] is null", e);
    g.out[e] = out[e];
  }
  if (op == MCM_LA_EIGVALS)
    for (int e = 0; e < nmat; ++e) {
      if (!out_im[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: out_im[%d] is null", e);
      g.out_im[e] = out_im[e];
    }
  if (n == 0) return MCM_OK;
  return dtype == MCM_F64 ? la_by_op<double>(op, nmat, g, m, n) : la_by_op<float>(op, nmat, g, m, n);
}

extern "C" int mcm_batched_linalg_complex(int op, int dtype, int nmat, const void* const* A_re, const void* const* A_im,
                                          void* const* out_re, void* const* out_im, size_t n) {
  MCM_NEED_INIT();
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: bad dtype %d", dtype);
  if (nmat < 1 || nmat > LA_MAXN) return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg_complex: matrix size must be 1..%d", LA_MAXN);
  if (!A_re || !out_re) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: null table");
  int nout = 0;
  switch (op) {
    case MCM_LAC_EXP: case MCM_LAC_LOG: nout = nmat * nmat; break;
    case MCM_LAC_DET: nout = 1; break;
    case MCM_LAC_EIGVALS: case MCM_LAC_SVDVALS: nout = nmat; break;
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: unknown operation %d", op);
  }
  if (op != MCM_LAC_SVDVALS && !out_im) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: this operation needs out_im");
  LacArgs g{};
  for (int e = 0; e < nmat * nmat; ++e) {
    if (!A_re[e] || (A_im && !A_im[e])) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: A[%d] is null", e);
    g.Are[e] = A_re[e];
    g.Aim[e] = A_im ? A_im[e] : nullptr;
  }
  for (int e = 0; e < nout; ++e) {
    if (!out_re[e] || (op != MCM_LAC_SVDVALS && !out_im[e])) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: out[%d] is null", e);
    g.ore[e] = out_re[e];
    g.oim[e] = op != MCM_LAC_SVDVALS ? out_im[e] : nullptr;
  }
  if (n == 0) return MCM_OK;
  return dtype == MCM_F64 ? lac_by_op<double>(op, nmat, g, A_im ? 0 : 1, n) : lac_by_op<float>(op, nmat, g, A_im ? 0 : 1, n);
}
