// complexops.cu — Complex{Particles}: elementwise complex arithmetic and functions (src/complex.jl:1-93).
// The reference maps `f(complex(re[i], im[i]))` over the particle index on the host (ℂ2ℂ_function, complex.jl:30-67)
// for ^ / sqrt exp exp10 log log10 sin cos tan; + - * come from Base's generic Complex arithmetic on the Particles
// fields.  Here a Complex{Particles} is two device vectors (re, im) and every operation is ONE streaming kernel:
// 16 B read + 16 B written per particle (32 + 16 for binary forms), the real and imaginary parts computed together.
// Arithmetic follows Base: * is (ac - bd) + (ad + bc)i with no contraction (bit-exact), + - componentwise; / and the
// functions are tolerance-tier restatements (scaled Smith division; log through log1p near |z| = 1; sqrt by the
// stable half-angle form; tan through tanh so that large imaginary parts cannot overflow).
#include <math.h>
#include "mcm_common.cuh"

namespace mcm {

struct CArg {            // one complex operand: each component is a device vector or a constant
  const void* re;
  const void* im;
  double sre, sim;
};
struct Cx { double x, y; };

template <typename T>
__device__ __forceinline__ Cx c_load(const CArg& a, size_t i) {
  Cx z;
  z.x = a.re ? (double)static_cast<const T*>(a.re)[i] : a.sre;
  z.y = a.im ? (double)static_cast<const T*>(a.im)[i] : a.sim;
  return z;
}

__device__ __forceinline__ Cx c_exp(Cx z) {
  const double e = exp(z.x);
  if (z.y == 0.0) return Cx{e, z.y};
  double s, c;
  sincos(z.y, &s, &c);
  return Cx{e * c, e * s};
}
__device__ __forceinline__ Cx c_log(Cx z) {
  const double ax = fmax(fabs(z.x), fabs(z.y)), ay = fmin(fabs(z.x), fabs(z.y));
  double re;
  if (ax > 0.5 && ax < 2.0) re = 0.5 * log1p(fma(ax - 1.0, ax + 1.0, ay * ay));   // |z|^2 - 1 without cancellation
  else re = log(hypot(z.x, z.y));
  return Cx{re, atan2(z.y, z.x)};
}
__device__ __forceinline__ Cx c_sqrt(Cx z) {
  if (z.x == 0.0 && z.y == 0.0) return Cx{0.0, z.y};
  const double t = sqrt(0.5 * (fabs(z.x) + hypot(z.x, z.y)));
  if (z.x >= 0.0) return Cx{t, z.y / (2.0 * t)};
  return Cx{fabs(z.y) / (2.0 * t), copysign(t, z.y)};
}
__device__ __forceinline__ Cx c_sin(Cx z) {
  double s, c;
  sincos(z.x, &s, &c);
  return Cx{s * cosh(z.y), c * sinh(z.y)};
}
__device__ __forceinline__ Cx c_cos(Cx z) {
  double s, c;
  sincos(z.x, &s, &c);
  return Cx{c * cosh(z.y), -(s * sinh(z.y))};
}
__device__ __forceinline__ Cx c_tan(Cx z) {
  // tan(x + iy) = (t (1 - h^2) + i h (1 + t^2)) / (1 + t^2 h^2), t = tan x, h = tanh y
  const double t = tan(z.x), h = tanh(z.y);
  const double den = 1.0 + (t * t) * (h * h);
  return Cx{t * (1.0 - h * h) / den, h * (1.0 + t * t) / den};
}
__device__ __forceinline__ Cx c_mul(Cx a, Cx b) { return Cx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ Cx c_div(Cx a, Cx b) {
  // Smith's algorithm; exact special cases of a real divisor kept (the reference's robust_cdiv agrees to rounding)
  if (b.y == 0.0) return Cx{a.x / b.x, a.y / b.x};
  if (fabs(b.x) >= fabs(b.y)) {
    const double r = b.y / b.x, d = b.x + b.y * r;
    return Cx{(a.x + a.y * r) / d, (a.y - a.x * r) / d};
  }
  const double r = b.x / b.y, d = b.x * r + b.y;
  return Cx{(a.x * r + a.y) / d, (a.y * r - a.x) / d};
}
__device__ __forceinline__ Cx c_pow(Cx z, Cx w) {
  if (z.x == 0.0 && z.y == 0.0) {
    if (w.x == 0.0 && w.y == 0.0) return Cx{1.0, 0.0};
    return w.x > 0.0 ? Cx{0.0, 0.0} : Cx{__builtin_huge_val(), 0.0};
  }
  if (w.y == 0.0 && z.y == 0.0 && z.x > 0.0) return Cx{pow(z.x, w.x), 0.0};
  return c_exp(c_mul(w, c_log(z)));
}

template <int OP, typename T>
__global__ void __launch_bounds__(kThreads)
k_complex(const __grid_constant__ CArg a, const __grid_constant__ CArg b, T* __restrict__ ore, T* __restrict__ oim, size_t n) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const Cx z = c_load<T>(a, i);
    Cx r;
    if constexpr (OP == MCM_C_ADD) { const Cx w = c_load<T>(b, i); r = Cx{z.x + w.x, z.y + w.y}; }
    else if constexpr (OP == MCM_C_SUB) { const Cx w = c_load<T>(b, i); r = Cx{z.x - w.x, z.y - w.y}; }
    else if constexpr (OP == MCM_C_MUL) r = c_mul(z, c_load<T>(b, i));
    else if constexpr (OP == MCM_C_DIV) r = c_div(z, c_load<T>(b, i));
    else if constexpr (OP == MCM_C_POW) r = c_pow(z, c_load<T>(b, i));
    else if constexpr (OP == MCM_C_EXP) r = c_exp(z);
    else if constexpr (OP == MCM_C_LOG) r = c_log(z);
    else if constexpr (OP == MCM_C_SQRT) r = c_sqrt(z);
    else if constexpr (OP == MCM_C_SIN) r = c_sin(z);
    else if constexpr (OP == MCM_C_COS) r = c_cos(z);
    else if constexpr (OP == MCM_C_TAN) r = c_tan(z);
    else if constexpr (OP == MCM_C_EXP10) r = c_exp(Cx{z.x * 2.302585092994046, z.y * 2.302585092994046});
    else if constexpr (OP == MCM_C_LOG10) { const Cx l = c_log(z); r = Cx{l.x * 0.4342944819032518, l.y * 0.4342944819032518}; }
    ore[i] = (T)r.x;
    oim[i] = (T)r.y;
  }
}

template <int OP>
static int complex_launch(int dtype, const CArg& a, const CArg& b, void* ore, void* oim, size_t n) {
  Ctx& c = ctx();
  const int grid = stream_grid(n, 1, 8);
  if (dtype == MCM_F64) k_complex<OP, double><<<grid, kThreads, 0, c.stream>>>(a, b, (double*)ore, (double*)oim, n);
  else k_complex<OP, float><<<grid, kThreads, 0, c.stream>>>(a, b, (float*)ore, (float*)oim, n);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

// z = op(a, b) or op(a): a = (a_re | sa_re) + i (a_im | sa_im) — a null component pointer selects the constant.
extern "C" int mcm_complex_op(int op, int dtype, const void* a_re, const void* a_im, double sa_re, double sa_im,
                              const void* b_re, const void* b_im, double sb_re, double sb_im, void* out_re, void* out_im,
                              size_t n) {
  MCM_NEED_INIT();
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_complex_op: bad dtype %d", dtype);
  if (n && (!out_re || !out_im)) return fail(MCM_ERR_INVALID, "mcm_complex_op: null output");
  if (n == 0) return MCM_OK;
  const CArg a{a_re, a_im, sa_re, sa_im}, b{b_re, b_im, sb_re, sb_im};
  switch (op) {
#define MCM_CCASE(OP) case OP: return complex_launch<OP>(dtype, a, b, out_re, out_im, n);
    MCM_CCASE(MCM_C_ADD) MCM_CCASE(MCM_C_SUB) MCM_CCASE(MCM_C_MUL) MCM_CCASE(MCM_C_DIV) MCM_CCASE(MCM_C_POW)
    MCM_CCASE(MCM_C_EXP) MCM_CCASE(MCM_C_LOG) MCM_CCASE(MCM_C_SQRT) MCM_CCASE(MCM_C_SIN) MCM_CCASE(MCM_C_COS)
    MCM_CCASE(MCM_C_TAN) MCM_CCASE(MCM_C_EXP10) MCM_CCASE(MCM_C_LOG10)
#undef MCM_CCASE
    default: return fail(MCM_ERR_INVALID, "mcm_complex_op: unknown operation %d", op);
  }
}
