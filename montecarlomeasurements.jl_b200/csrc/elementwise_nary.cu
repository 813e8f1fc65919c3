// elementwise_nary.cu — f(p1,p2), f(p,a::Real), f(a::Real,p) and the three-operand
// forms (clamp, muladd) of register_primitive_multi (src/particles.jl:183,
// bodies src/register_primitive.jl:44-58; muladd src/particles.jl:416-422;
// rem/div src/particles.jl:354-358), plus the fused affine / polynomial /
// linear-combination chains.
#include "mcm_map.cuh"
#include "mcm_math.cuh"

namespace mcm {

template <int OP>
struct FBinary {
  template <typename T>
  __device__ __forceinline__ T operator()(T x, T y) const { return math::binary_apply<OP, T>(x, y); }
};
template <int OP>
struct FTernary {
  template <typename T>
  __device__ __forceinline__ T operator()(T x, T y, T z) const { return math::ternary_apply<OP, T>(x, y, z); }
};

template <int OP>
constexpr int binary_unroll() {
  return (OP == MCM_B_ADD || OP == MCM_B_SUB || OP == MCM_B_MUL || OP == MCM_B_ADD_SUM ||
          OP == MCM_B_MAX || OP == MCM_B_MIN)
             ? 4
             : 2;
}

template <int OP, typename T>
static int run_binary(const void* a, double sa, const void* b, double sb, void* out, size_t n) {
  Operands<T, 2> ops;
  ops.p[0] = static_cast<const T*>(a); ops.s[0] = (T)sa;
  ops.p[1] = static_cast<const T*>(b); ops.s[1] = (T)sb;
  return launch_map<T, 2, binary_unroll<OP>(), math::BinaryCanDomainError<OP>::value>(
      ops, static_cast<T*>(out), n, FBinary<OP>());
}

template <typename T>
static int dispatch_binary(int op, const void* a, double sa, const void* b, double sb, void* out, size_t n) {
  switch (op) {
#define MCM_CASE(OP) \
  case OP: return run_binary<OP, T>(a, sa, b, sb, out, n);
    MCM_CASE(MCM_B_ADD) MCM_CASE(MCM_B_SUB) MCM_CASE(MCM_B_MUL) MCM_CASE(MCM_B_DIV)
    MCM_CASE(MCM_B_POW) MCM_CASE(MCM_B_MAX) MCM_CASE(MCM_B_MIN) MCM_CASE(MCM_B_MOD)
    MCM_CASE(MCM_B_MOD1) MCM_CASE(MCM_B_ATAN2) MCM_CASE(MCM_B_ATAND2) MCM_CASE(MCM_B_ADD_SUM)
    MCM_CASE(MCM_B_HYPOT) MCM_CASE(MCM_B_REM) MCM_CASE(MCM_B_FDIV)
    MCM_CASE(MCM_B_REM_NEAREST) MCM_CASE(MCM_B_REM_DOWN) MCM_CASE(MCM_B_REM_UP) MCM_CASE(MCM_B_REM_FROMZERO)
    MCM_CASE(MCM_B_DIV_NEAREST) MCM_CASE(MCM_B_DIV_DOWN) MCM_CASE(MCM_B_DIV_UP) MCM_CASE(MCM_B_DIV_FROMZERO)
#undef MCM_CASE
    case MCM_B_RATDIV:
      return fail(MCM_ERR_UNSUPPORTED, "mcm_binary: `//` has no floating-point method (MethodError in Base)");
    default: return fail(MCM_ERR_INVALID, "mcm_binary: unknown op %d", op);
  }
}

template <int OP, typename T>
static int run_ternary(const void* a, double sa, const void* b, double sb, const void* c, double sc,
                       void* out, size_t n) {
  Operands<T, 3> ops;
  ops.p[0] = static_cast<const T*>(a); ops.s[0] = (T)sa;
  ops.p[1] = static_cast<const T*>(b); ops.s[1] = (T)sb;
  ops.p[2] = static_cast<const T*>(c); ops.s[2] = (T)sc;
  return launch_map<T, 3, 2, false>(ops, static_cast<T*>(out), n, FTernary<OP>());
}

template <typename T>
static int dispatch_ternary(int op, const void* a, double sa, const void* b, double sb, const void* c,
                            double sc, void* out, size_t n) {
  switch (op) {
    case MCM_T_CLAMP: return run_ternary<MCM_T_CLAMP, T>(a, sa, b, sb, c, sc, out, n);
    case MCM_T_MULADD: return run_ternary<MCM_T_MULADD, T>(a, sa, b, sb, c, sc, out, n);
    case MCM_T_MUL_ADD: return run_ternary<MCM_T_MUL_ADD, T>(a, sa, b, sb, c, sc, out, n);
    default: return fail(MCM_ERR_INVALID, "mcm_ternary: unknown op %d", op);
  }
}

// ---- fused chains -------------------------------------------------------------
static constexpr int kMaxPoly = 32;
template <typename T, bool FUSED>
struct FPoly {
  T c[kMaxPoly];
  int nc;
  __device__ __forceinline__ T operator()(T x) const {
    T acc = c[nc - 1];
    for (int k = nc - 2; k >= 0; --k) {
      if constexpr (FUSED) acc = math::m_fma(acc, x, c[k]);
      else acc = acc * x + c[k];
    }
    return acc;
  }
};

// degree known at compile time (2..16 coefficients): the Horner recurrences of the values a thread holds unroll and
// interleave, and the coefficients are constant-bank operands — with the run-time trip count above every step is a
// dependent FMA behind an indexed constant load (poly8: 0.84-0.90 of the HBM rate)
template <typename T, bool FUSED, int NC>
struct FPolyN {
  T c[NC];
  __device__ __forceinline__ T operator()(T x) const {
    T acc = c[NC - 1];
#pragma unroll
    for (int k = NC - 2; k >= 0; --k) {
      if constexpr (FUSED) acc = math::m_fma(acc, x, c[k]);
      else acc = acc * x + c[k];
    }
    return acc;
  }
};

template <typename T>
struct FAffine {
  T alpha, beta;
  __device__ __forceinline__ T operator()(T x) const { return alpha * x + beta; }
};

static constexpr int kMaxTerms = 8;
// out = c0*x0 + c1*x1 + ... evaluated left to right with separate roundings
// (bit-identical to the chain of reference broadcasts `c0*x0 + c1*x1 + ...`).
template <typename T, int NT>
__global__ void __launch_bounds__(kThreads)
k_lincomb(const T* const* __restrict__ xs_unused, Operands<T, NT> ops, T* __restrict__ out, size_t n) {
  using P = Pack<T>;
  constexpr int V = P::N;
  const size_t nvec = n / V;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  for (size_t i = tid; i < nvec; i += stride * 2) {
    P x[NT][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t j = i + (size_t)u * stride;
      if (j < nvec) {
#pragma unroll
        for (int k = 0; k < NT; ++k) x[k][u] = ld_stream(reinterpret_cast<const P*>(ops.p[k]) + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t j = i + (size_t)u * stride;
      if (j < nvec) {
        P r;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          T acc = ops.s[0] * x[0][u].v[e];
#pragma unroll
          for (int k = 1; k < NT; ++k) acc = acc + ops.s[k] * x[k][u].v[e];
          r.v[e] = acc;
        }
        st_stream(reinterpret_cast<P*>(out) + j, r);
      }
    }
  }
  for (size_t i = nvec * V + tid; i < n; i += stride) {
    T acc = ops.s[0] * ops.p[0][i];
#pragma unroll
    for (int k = 1; k < NT; ++k) acc = acc + ops.s[k] * ops.p[k][i];
    out[i] = acc;
  }
}

template <typename T, int NT>
static int run_lincomb(const void* const* xs, const double* coeffs, void* out, size_t n) {
  Operands<T, NT> ops;
  bool vec = aligned16(out);
  for (int k = 0; k < NT; ++k) {
    ops.p[k] = static_cast<const T*>(xs[k]);
    ops.s[k] = (T)coeffs[k];
    vec = vec && aligned16(xs[k]);
    if (!xs[k]) return fail(MCM_ERR_INVALID, "mcm_lincomb: null term %d", k);
  }
  if (!vec) return fail(MCM_ERR_INVALID, "mcm_lincomb: pointers must be 16-byte aligned");
  Ctx& c = ctx();
  const int grid = stream_grid((n + Pack<T>::N - 1) / Pack<T>::N, 2, 6);
  k_lincomb<T, NT><<<grid, kThreads, 0, c.stream>>>(nullptr, ops, static_cast<T*>(out), n);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

template <typename T>
static int dispatch_lincomb(int nt, const void* const* xs, const double* coeffs, void* out, size_t n) {
  switch (nt) {
    case 1: return run_lincomb<T, 1>(xs, coeffs, out, n);
    case 2: return run_lincomb<T, 2>(xs, coeffs, out, n);
    case 3: return run_lincomb<T, 3>(xs, coeffs, out, n);
    case 4: return run_lincomb<T, 4>(xs, coeffs, out, n);
    case 5: return run_lincomb<T, 5>(xs, coeffs, out, n);
    case 6: return run_lincomb<T, 6>(xs, coeffs, out, n);
    case 7: return run_lincomb<T, 7>(xs, coeffs, out, n);
    case 8: return run_lincomb<T, 8>(xs, coeffs, out, n);
    default: return fail(MCM_ERR_INVALID, "mcm_lincomb: nterms must be 1..%d", kMaxTerms);
  }
}

template <typename T, bool FUSED, int NC>
static int run_poly_n(const Operands<T, 1>& ops, const double* coeffs, void* out, size_t n) {
  FPolyN<T, FUSED, NC> f;
  for (int k = 0; k < NC; ++k) f.c[k] = (T)coeffs[k];
  return launch_map<T, 1, 4, false>(ops, static_cast<T*>(out), n, f);
}
template <typename T>
static int run_poly(const void* in, const double* coeffs, int nc, int fused, void* out, size_t n) {
  Operands<T, 1> ops;
  ops.p[0] = static_cast<const T*>(in);
  ops.s[0] = 0;
  switch (nc) {
#define MCM_POLY_CASE(NC) \
  case NC: return fused ? run_poly_n<T, true, NC>(ops, coeffs, out, n) : run_poly_n<T, false, NC>(ops, coeffs, out, n);
    MCM_POLY_CASE(2) MCM_POLY_CASE(3) MCM_POLY_CASE(4) MCM_POLY_CASE(5) MCM_POLY_CASE(6) MCM_POLY_CASE(7) MCM_POLY_CASE(8)
    MCM_POLY_CASE(9) MCM_POLY_CASE(10) MCM_POLY_CASE(11) MCM_POLY_CASE(12) MCM_POLY_CASE(13) MCM_POLY_CASE(14)
    MCM_POLY_CASE(15) MCM_POLY_CASE(16)
#undef MCM_POLY_CASE
    default: break;
  }
  if (fused) {
    FPoly<T, true> f;
    for (int k = 0; k < kMaxPoly; ++k) f.c[k] = k < nc ? (T)coeffs[k] : T(0);
    f.nc = nc;
    return launch_map<T, 1, 4, false>(ops, static_cast<T*>(out), n, f);
  }
  FPoly<T, false> f;
  for (int k = 0; k < kMaxPoly; ++k) f.c[k] = k < nc ? (T)coeffs[k] : T(0);
  f.nc = nc;
  return launch_map<T, 1, 4, false>(ops, static_cast<T*>(out), n, f);
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_binary(int op, int dtype, const void* a, double sa, const void* b, double sb,
                          void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_binary: null output");
  if (!a && !b) return fail(MCM_ERR_INVALID, "mcm_binary: at least one operand must be particles");
  if (dtype == MCM_F64) return dispatch_binary<double>(op, a, sa, b, sb, out, n);
  if (dtype == MCM_F32) return dispatch_binary<float>(op, a, sa, b, sb, out, n);
  return fail(MCM_ERR_INVALID, "mcm_binary: bad dtype %d", dtype);
}

extern "C" int mcm_ternary(int op, int dtype, const void* a, double sa, const void* b, double sb,
                           const void* c, double sc, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_ternary: null output");
  if (!a && !b && !c) return fail(MCM_ERR_INVALID, "mcm_ternary: at least one operand must be particles");
  if (dtype == MCM_F64) return dispatch_ternary<double>(op, a, sa, b, sb, c, sc, out, n);
  if (dtype == MCM_F32) return dispatch_ternary<float>(op, a, sa, b, sb, c, sc, out, n);
  return fail(MCM_ERR_INVALID, "mcm_ternary: bad dtype %d", dtype);
}

extern "C" int mcm_polyval(int dtype, const void* in, const double* coeffs_host, int ncoef, int fused,
                           void* out, size_t n) {
  MCM_NEED_INIT();
  if (ncoef < 1 || ncoef > kMaxPoly) return fail(MCM_ERR_INVALID, "mcm_polyval: ncoef must be 1..%d", kMaxPoly);
  if (!coeffs_host || (n && (!in || !out))) return fail(MCM_ERR_INVALID, "mcm_polyval: null pointer");
  if (dtype == MCM_F64) return run_poly<double>(in, coeffs_host, ncoef, fused, out, n);
  if (dtype == MCM_F32) return run_poly<float>(in, coeffs_host, ncoef, fused, out, n);
  return fail(MCM_ERR_INVALID, "mcm_polyval: bad dtype %d", dtype);
}

extern "C" int mcm_affine(int dtype, const void* in, double alpha, double beta, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_affine: null pointer");
  if (dtype == MCM_F64) {
    Operands<double, 1> ops; ops.p[0] = (const double*)in; ops.s[0] = 0;
    return launch_map<double, 1, 4, false>(ops, (double*)out, n, FAffine<double>{alpha, beta});
  }
  if (dtype == MCM_F32) {
    Operands<float, 1> ops; ops.p[0] = (const float*)in; ops.s[0] = 0;
    return launch_map<float, 1, 4, false>(ops, (float*)out, n, FAffine<float>{(float)alpha, (float)beta});
  }
  return fail(MCM_ERR_INVALID, "mcm_affine: bad dtype %d", dtype);
}

extern "C" int mcm_lincomb(int dtype, int nterms, const void* const* xs, const double* coeffs_host,
                           void* out, size_t n) {
  MCM_NEED_INIT();
  if (!xs || !coeffs_host || (n && !out)) return fail(MCM_ERR_INVALID, "mcm_lincomb: null pointer");
  if (n == 0) return MCM_OK;
  if (dtype == MCM_F64) return dispatch_lincomb<double>(nterms, xs, coeffs_host, out, n);
  if (dtype == MCM_F32) return dispatch_lincomb<float>(nterms, xs, coeffs_host, out, n);
  return fail(MCM_ERR_INVALID, "mcm_lincomb: bad dtype %d", dtype);
}
