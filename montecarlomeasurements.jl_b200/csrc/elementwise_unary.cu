// elementwise_unary.cu — f(p) for the 35 register_primitive_single entries
// (src/particles.jl:185-190, bodies at src/register_primitive.jl:94-97 and
// src/sleefpirates.jl:7-16,26-35) plus round/floor/ceil/trunc/inv/sqr/cube
// (src/particles.jl:346-348,388-391), fill and convert (types.jl:57-63,
// particles.jl:333).
#include "mcm_map.cuh"
#include "mcm_math.cuh"

namespace mcm {

template <int OP>
struct FUnary {
  template <typename T>
  __device__ __forceinline__ T operator()(T x) const { return math::unary_apply<OP, T>(x); }
};

// cheap exact-tier ops stream with 4 requests in flight, transcendental ones with 2
template <int OP>
constexpr int unary_unroll() {
  return (OP == MCM_U_PLUS || OP == MCM_U_NEG || OP == MCM_U_ZERO || OP == MCM_U_SIGN ||
          OP == MCM_U_ABS || OP == MCM_U_RAD2DEG || OP == MCM_U_DEG2RAD || OP == MCM_U_FLOAT ||
          OP == MCM_U_ROUND || OP == MCM_U_FLOOR || OP == MCM_U_CEIL || OP == MCM_U_TRUNC ||
          OP == MCM_U_SQR || OP == MCM_U_CUBE || OP == MCM_U_ROUND_TIESAWAY || OP == MCM_U_ROUND_TIESUP ||
          OP == MCM_U_ROUND_FROMZERO)
             ? 4
             : 2;
}

template <int OP, typename T>
static int run_unary(const void* in, void* out, size_t n) {
  Operands<T, 1> ops;
  ops.p[0] = static_cast<const T*>(in);
  ops.s[0] = T(0);
  return launch_map<T, 1, unary_unroll<OP>(), math::UnaryCanDomainError<OP>::value>(
      ops, static_cast<T*>(out), n, FUnary<OP>());
}

template <typename T>
static int dispatch_unary(int op, const void* in, void* out, size_t n) {
  switch (op) {
#define MCM_CASE(OP) \
  case OP: return run_unary<OP, T>(in, out, n);
    MCM_CASE(MCM_U_PLUS) MCM_CASE(MCM_U_NEG) MCM_CASE(MCM_U_EXP) MCM_CASE(MCM_U_EXP2)
    MCM_CASE(MCM_U_EXP10) MCM_CASE(MCM_U_EXPM1) MCM_CASE(MCM_U_LOG) MCM_CASE(MCM_U_LOG10)
    MCM_CASE(MCM_U_LOG2) MCM_CASE(MCM_U_LOG1P) MCM_CASE(MCM_U_SIN) MCM_CASE(MCM_U_COS)
    MCM_CASE(MCM_U_TAN) MCM_CASE(MCM_U_SIND) MCM_CASE(MCM_U_COSD) MCM_CASE(MCM_U_TAND)
    MCM_CASE(MCM_U_SINH) MCM_CASE(MCM_U_COSH) MCM_CASE(MCM_U_TANH) MCM_CASE(MCM_U_ASIN)
    MCM_CASE(MCM_U_ACOS) MCM_CASE(MCM_U_ATAN) MCM_CASE(MCM_U_ASIND) MCM_CASE(MCM_U_ACOSD)
    MCM_CASE(MCM_U_ATAND) MCM_CASE(MCM_U_ASINH) MCM_CASE(MCM_U_ACOSH) MCM_CASE(MCM_U_ATANH)
    MCM_CASE(MCM_U_ZERO) MCM_CASE(MCM_U_SIGN) MCM_CASE(MCM_U_ABS) MCM_CASE(MCM_U_SQRT)
    MCM_CASE(MCM_U_RAD2DEG) MCM_CASE(MCM_U_DEG2RAD) MCM_CASE(MCM_U_FLOAT) MCM_CASE(MCM_U_ROUND)
    MCM_CASE(MCM_U_FLOOR) MCM_CASE(MCM_U_CEIL) MCM_CASE(MCM_U_TRUNC) MCM_CASE(MCM_U_INV)
    MCM_CASE(MCM_U_SQR) MCM_CASE(MCM_U_CUBE)
    MCM_CASE(MCM_U_ROUND_TIESAWAY) MCM_CASE(MCM_U_ROUND_TIESUP) MCM_CASE(MCM_U_ROUND_FROMZERO)
    MCM_CASE(MCM_U_REM2PI_NEAREST) MCM_CASE(MCM_U_REM2PI_TOZERO) MCM_CASE(MCM_U_REM2PI_DOWN) MCM_CASE(MCM_U_REM2PI_UP)
#undef MCM_CASE
    default: return fail(MCM_ERR_INVALID, "mcm_unary: unknown op %d", op);
  }
}

struct FPowi {
  int k;
  template <typename T>
  __device__ __forceinline__ T operator()(T x) const { return math::m_powi(x, k); }
};

template <typename T>
struct FFill {
  T v;
  __device__ __forceinline__ T operator()(T) const { return v; }
};

// convert between element types (one pack in, one pack out would need two pack
// widths; a scalar grid-stride loop with 8-byte/4-byte accesses is still coalesced)
template <typename S, typename D>
__global__ void __launch_bounds__(kThreads) k_convert(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) dst[i] = (D)src[i];
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_unary(int op, int dtype, const void* in, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_unary: null pointer");
  if (dtype == MCM_F64) return dispatch_unary<double>(op, in, out, n);
  if (dtype == MCM_F32) return dispatch_unary<float>(op, in, out, n);
  return fail(MCM_ERR_INVALID, "mcm_unary: bad dtype %d", dtype);
}

extern "C" int mcm_powi(int dtype, const void* in, int k, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_powi: null pointer");
  FPowi f{k};
  if (dtype == MCM_F64) {
    Operands<double, 1> ops; ops.p[0] = (const double*)in; ops.s[0] = 0;
    return launch_map<double, 1, 2, false>(ops, (double*)out, n, f);
  }
  if (dtype == MCM_F32) {
    Operands<float, 1> ops; ops.p[0] = (const float*)in; ops.s[0] = 0;
    return launch_map<float, 1, 2, false>(ops, (float*)out, n, f);
  }
  return fail(MCM_ERR_INVALID, "mcm_powi: bad dtype %d", dtype);
}

extern "C" int mcm_fill(int dtype, void* out, size_t n, double value) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_fill: null pointer");
  if (dtype == MCM_F64) {
    Operands<double, 1> ops; ops.p[0] = nullptr; ops.s[0] = value;
    return launch_map<double, 1, 4, false>(ops, (double*)out, n, FFill<double>{value});
  }
  if (dtype == MCM_F32) {
    Operands<float, 1> ops; ops.p[0] = nullptr; ops.s[0] = (float)value;
    return launch_map<float, 1, 4, false>(ops, (float*)out, n, FFill<float>{(float)value});
  }
  return fail(MCM_ERR_INVALID, "mcm_fill: bad dtype %d", dtype);
}

extern "C" int mcm_convert(int src_dtype, const void* src, int dst_dtype, void* dst, size_t n) {
  MCM_NEED_INIT();
  if (n == 0) return MCM_OK;
  if (!src || !dst) return fail(MCM_ERR_INVALID, "mcm_convert: null pointer");
  Ctx& c = ctx();
  const int grid = stream_grid(n, 4, 8);
  if (src_dtype == MCM_F32 && dst_dtype == MCM_F64)
    k_convert<float, double><<<grid, kThreads, 0, c.stream>>>((const float*)src, (double*)dst, n);
  else if (src_dtype == MCM_F64 && dst_dtype == MCM_F32)
    k_convert<double, float><<<grid, kThreads, 0, c.stream>>>((const double*)src, (float*)dst, n);
  else if (src_dtype == dst_dtype && (src_dtype == MCM_F32 || src_dtype == MCM_F64)) {
    MCM_CUDA(cudaMemcpyAsync(dst, src, n * (src_dtype == MCM_F64 ? 8 : 4), cudaMemcpyDeviceToDevice, c.stream));
    return MCM_OK;
  } else
    return fail(MCM_ERR_INVALID, "mcm_convert: bad dtypes %d -> %d", src_dtype, dst_dtype);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
