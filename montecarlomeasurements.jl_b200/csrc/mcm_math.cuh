// mcm_math.cuh — per-particle scalar semantics of every registered primitive.
//
// Each function states which reference method body it stands for
// (src/particles.jl:183-190 lists the primitives; src/sleefpirates.jl:2-36 says
// which of them are SLEEFPirates calls for Float32/Float64).  Exact-tier ops
// (+ - * / neg abs sqrt min max sign round rem mod clamp literal powers) are
// IEEE-754 correctly rounded and the library is built with -fmad=false, so they
// are bit-identical to Julia's broadcast.  Transcendentals are ulp-tier: see
// DESIGN.md §accuracy for the measured bounds against the oracle.
#pragma once
#include <math.h>
#ifdef __CUDACC_RTC__
#include "mcm_b200.h"
#else
#include "../../include/mcm_b200.h"
#endif
#include "mcm_fastmath.cuh"

namespace mcm {
namespace math {

#define MCM_DI __device__ __forceinline__

template <typename T> struct C;
template <> struct C<double> {
  static constexpr double rad2deg = 180.0 / 3.141592653589793;  // Base: z * (180 / oftype(z, pi))
  static constexpr double deg2rad = 3.141592653589793 / 180.0;  // Base: z * (oftype(z, pi) / 180)
  static constexpr double d2r_hi = 0.017453292519943295;        // pi/180 as double-double
  static constexpr double d2r_lo = 2.9486522708701687e-19;
};
template <> struct C<float> {
  static constexpr float rad2deg = 180.0f / 3.14159274f;
  static constexpr float deg2rad = 3.14159274f / 180.0f;
  static constexpr float d2r_hi = 0.01745329238474369f;
  static constexpr float d2r_lo = 1.3519960151420207e-10f;
};

// ---- thin overload layer so the op table below is type-generic -------------
MCM_DI double m_fma(double a, double b, double c) { return ::fma(a, b, c); }
MCM_DI float m_fma(float a, float b, float c) { return ::fmaf(a, b, c); }
MCM_DI double m_rint(double x) { return ::rint(x); }
MCM_DI float m_rint(float x) { return ::rintf(x); }
MCM_DI double m_fmod(double x, double y) { return ::fmod(x, y); }
MCM_DI float m_fmod(float x, float y) { return ::fmodf(x, y); }
MCM_DI double m_copysign(double x, double y) { return ::copysign(x, y); }
MCM_DI float m_copysign(float x, float y) { return ::copysignf(x, y); }
MCM_DI double m_abs(double x) { return ::fabs(x); }
MCM_DI float m_abs(float x) { return ::fabsf(x); }

#define MCM_FWD1(name, fd, ff)                      \
  MCM_DI double name(double x) { return fd(x); }    \
  MCM_DI float name(float x) { return ff(x); }
// SLEEF-backed in the reference (sleefpirates.jl): exp, log, sin, cos, asin, acos, atan.
// The double versions of the hot ones are our own kernels (mcm_fastmath.cuh).
MCM_DI double m_exp(double x) { return fast::exp_f64(x); }
MCM_DI float m_exp(float x) { return ::expf(x); }
MCM_DI double m_sin(double x) { return fast::sin_f64(x); }
MCM_DI float m_sin(float x) { return ::sinf(x); }
MCM_DI double m_cos(double x) { return fast::cos_f64(x); }
MCM_DI float m_cos(float x) { return ::cosf(x); }
MCM_DI double m_log(double x) { return fast::log_f64(x); }
MCM_DI float m_log(float x) { return ::logf(x); }
MCM_FWD1(m_asin, ::asin, ::asinf)
MCM_FWD1(m_acos, ::acos, ::acosf)
MCM_FWD1(m_atan, ::atan, ::atanf)
// Base Julia scalar math in the reference
MCM_FWD1(m_exp2, ::exp2, ::exp2f)
MCM_FWD1(m_exp10, ::exp10, ::exp10f)
MCM_FWD1(m_expm1, ::expm1, ::expm1f)
MCM_FWD1(m_log10, ::log10, ::log10f)
MCM_FWD1(m_log2, ::log2, ::log2f)
MCM_FWD1(m_log1p, ::log1p, ::log1pf)
// Float32: the CUDA float kernels of these are 2-4 ulp (tanf 4, sinhf 3, acoshf 4, atan2f 3 ...; measured against
// the truth on a B200: tan/sinh/tanh 2, vs a glibc oracle 3).  The reference's Float32 methods are < 1 ulp (Base), so
// the Float32 versions evaluate in Float64 and round once: <= 0.5000001 ulp.  These are not on the headline path.
#define MCM_FWD1_VIA64(name, fd)                       \
  MCM_DI double name(double x) { return fd(x); }      \
  MCM_DI float name(float x) { return (float)fd((double)x); }
MCM_FWD1_VIA64(m_tan, ::tan)
MCM_FWD1_VIA64(m_sinh, ::sinh)
MCM_FWD1_VIA64(m_cosh, ::cosh)
MCM_FWD1_VIA64(m_tanh, ::tanh)
MCM_FWD1_VIA64(m_asinh, ::asinh)
MCM_FWD1_VIA64(m_acosh, ::acosh)
MCM_FWD1_VIA64(m_atanh, ::atanh)
#undef MCM_FWD1_VIA64
MCM_FWD1(m_sqrt, ::sqrt, ::sqrtf)
MCM_FWD1(m_floor, ::floor, ::floorf)
MCM_FWD1(m_ceil, ::ceil, ::ceilf)
MCM_FWD1(m_trunc, ::trunc, ::truncf)
#undef MCM_FWD1
MCM_DI double m_pow(double x, double y) { return fast::pow_f64(x, y); }
MCM_DI float m_pow(float x, float y) { return ::powf(x, y); }
MCM_DI double m_atan2(double y, double x) { return ::atan2(y, x); }
MCM_DI float m_atan2(float y, float x) { return (float)::atan2((double)y, (double)x); }
// hypot: x^2 + y^2 as a double-double (exact products, Fast2Sum), correctly rounded sqrt of the high word and one
// Newton correction with the low word: <= 0.5 ulp + 2^-52 (the library version measured 2 ulp).  Scaling by a power
// of two keeps the squares in range; Inf wins over NaN like Base.hypot / C99.
MCM_DI double m_hypot(double x, double y) {
  double a = ::fabs(x), b = ::fabs(y);
  if (a == __builtin_huge_val() || b == __builtin_huge_val()) return __builtin_huge_val();
  if (a != a || b != b) return a + b;
  if (a < b) { const double t = a; a = b; b = t; }
  if (a == 0.0) return 0.0;
  double sc = 1.0, isc = 1.0;                                   // exact power-of-two scaling away from over/underflow
  if (a > 0x1p+500) { sc = 0x1p-600; isc = 0x1p+600; }
  else if (a < 0x1p-500) { sc = 0x1p+600; isc = 0x1p-600; }
  a *= sc; b *= sc;
  const double ah = a * a, al = ::fma(a, a, -ah);
  const double bh = b * b, bl = ::fma(b, b, -bh);
  const double sh = ah + bh;
  const double sl = (bh - (sh - ah)) + (al + bl);               // ah >= bh: Fast2Sum
  const double r = ::sqrt(sh);
  const double corr = (::fma(-r, r, sh) + sl) / (2.0 * r);
  return (r + corr) * isc;
}
MCM_DI float m_hypot(float x, float y) { return (float)::sqrt((double)x * (double)x + (double)y * (double)y); }
MCM_DI void m_sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
MCM_DI void m_sincos(float x, float* s, float* c) { ::sincosf(x, s, c); }

// sign(x): Base `ifelse(x < 0, -1, ifelse(x > 0, 1, x))` (NaN and +-0 pass through)
template <typename T> MCM_DI T m_sign(T x) { return x < T(0) ? T(-1) : (x > T(0) ? T(1) : x); }

// Base max/min for Float32/Float64: NaN-propagating, -0.0 < +0.0
template <typename T> MCM_DI T m_max(T x, T y) {
  T d = x - y;
  T arg = signbit(d) ? y : x;
  return (x != x || y != y) ? d : arg;
}
template <typename T> MCM_DI T m_min(T x, T y) {
  T d = x - y;
  T arg = signbit(d) ? x : y;
  return (x != x || y != y) ? d : arg;
}
// Base mod(x,y) for floats: floored modulo built on rem (= fmod)
template <typename T> MCM_DI T m_mod(T x, T y) {
  T r = m_fmod(x, y);
  if (r == T(0)) return m_copysign(r, y);
  if ((r > T(0)) != (y > T(0))) return r + y;
  return r;
}
template <typename T> MCM_DI T m_mod1(T x, T y) {
  T m = m_mod(x, y);
  return m == T(0) ? y : m;
}
// Base div(x,y) for floats: round((x - rem(x,y))/y)
template <typename T> MCM_DI T m_fdiv(T x, T y) { return m_rint((x - m_fmod(x, y)) / y); }
template <typename T> MCM_DI T m_clamp(T x, T lo, T hi) { return x > hi ? hi : (x < lo ? lo : x); }

// round(x, r::RoundingMode) beyond Nearest/ToZero/Up/Down (Base floatfuncs.jl / rounding.jl)
MCM_DI double m_eps_half(double) { return 1.1102230246251565e-16; }   // eps(0.5)
MCM_DI float m_eps_half(float) { return 5.9604645e-08f; }
template <typename T> MCM_DI T m_round_tiesaway(T x) {   // y = trunc(x); ifelse(x == y, y, trunc(2x - y))
  const T y = m_trunc(x);
  return x == y ? y : m_trunc(T(2) * x - y);
}
template <typename T> MCM_DI T m_round_tiesup(T x) {     // copysign(floor((x + (1/4 - eps(1/2))) + (1/4 + eps(1/2))), x)
  const T e = m_eps_half(x);
  return m_copysign(m_floor((x + (T(0.25) - e)) + (T(0.25) + e)), x);
}
template <typename T> MCM_DI T m_round_fromzero(T x) { return signbit(x) ? m_floor(x) : m_ceil(x); }

// rem(x, y, r) for the RoundingModes of Base div.jl; rem(x, y, RoundToZero) is m_fmod
MCM_DI double m_remainder(double x, double y) { return ::remainder(x, y); }
MCM_DI float m_remainder(float x, float y) { return ::remainderf(x, y); }
template <typename T> MCM_DI T m_rem_up(T x, T y) { return m_mod(x, -y); }
template <typename T> MCM_DI T m_rem_fromzero(T x, T y) { return signbit(x) == signbit(y) ? m_rem_up(x, y) : m_mod(x, y); }
// div(x, y, r) = round((x - rem(x, y, r)) / y)
template <typename T> MCM_DI T m_div_r(T x, T y, T r) { return m_rint((x - r) / y); }

// rem2pi(x, r): Float32 goes through Float64 like Base (rem2pi(x::Float32, r) = Float32(rem2pi(Float64(x), r)))
MCM_DI double m_rem2pi(double x, int mode) { return fast::rem2pi_f64(x, mode); }
MCM_DI float m_rem2pi(float x, int mode) { return (float)fast::rem2pi_f64((double)x, mode); }

// Degree-argument trig, following Base's sind/cosd structure: exact reduction
// of the argument in DEGREES (rem(x,360), then to |t|<=45 by exact subtractions),
// conversion to radians in double-double, kernel on the reduced argument.
// Exact results at multiples of 30/45/90 degrees come out as in Base.
template <typename T> MCM_DI void sincosd_core(T x, T* sd, T* cd) {
  if (!(m_abs(x) <= T(1.7976931348623157e308))) {  // Inf (DomainError in Base) or NaN
    *sd = *cd = x - x;
    return;
  }
  T rx = m_fmod(x, T(360));  // exact; sign of x
  T arx = m_abs(rx);
  // quadrant q in 0..3 and t in [-45,45] with arx = 90 q + t (all subtractions exact)
  int q;
  T t;
  if (arx < T(45)) { q = 0; t = arx; }
  else if (arx <= T(135)) { q = 1; t = arx - T(90); }
  else if (arx < T(225)) { q = 2; t = arx - T(180); }
  else if (arx <= T(315)) { q = 3; t = arx - T(270); }
  else { q = 0; t = arx - T(360); }
  T hi = t * C<T>::d2r_hi;
  T lo = m_fma(t, C<T>::d2r_hi, -hi) + t * C<T>::d2r_lo;
  T s, c;
  m_sincos(hi, &s, &c);
  T s2 = m_fma(lo, c, s);   // sin(hi+lo)
  T c2 = m_fma(-lo, s, c);  // cos(hi+lo)
  if (t == T(0)) { s2 = T(0); c2 = T(1); }
  T sa, ca;  // sin/cos of |rx|
  switch (q) {
    case 0: sa = s2; ca = c2; break;
    case 1: sa = c2; ca = -s2; break;
    case 2: sa = -s2; ca = -c2; break;
    default: sa = -c2; ca = s2; break;
  }
  // Base returns +0 for the cosine zeros and a zero carrying the sign of x for sine zeros
  if (ca == T(0)) ca = T(0);
  if (sa == T(0)) sa = T(0);
  *sd = (rx == T(0)) ? rx : m_copysign(T(1), rx) * sa;
  if (rx != T(0) && sa == T(0)) *sd = m_copysign(T(0), rx);
  *cd = ca;
}
template <typename T> MCM_DI T m_sind(T x) { T s, c; sincosd_core(x, &s, &c); return s; }
template <typename T> MCM_DI T m_cosd(T x) { T s, c; sincosd_core(x, &s, &c); return c; }
template <typename T> MCM_DI T m_tand(T x) { T s, c; sincosd_core(x, &s, &c); return s / c; }

// p^k for integer k (particles.jl:383-391).  k in {-2,...,3} is literally what
// Base.literal_pow expands to (x*x, x*x*x, inv(x), inv(x)^2) and is bit-exact.
// Other k: compensated power-by-squaring (Base.Math.pow_body structure), which
// is within 1 ulp of the exact power.
MCM_DI void two_mul(double a, double b, double* hi, double* lo) { *hi = a * b; *lo = ::fma(a, b, -*hi); }
MCM_DI void two_mul(float a, float b, float* hi, float* lo) { *hi = a * b; *lo = ::fmaf(a, b, -*hi); }
template <typename T> MCM_DI T m_powi(T x, int n) {
  if (n == 0) return T(1);
  if (n == 1) return x;
  if (n == 2) return x * x;
  if (n == 3) return x * x * x;
  if (n == -1) return T(1) / x;
  T y = T(1), xnlo = T(0), ynlo = T(0);
  if (n < 0) {
    T rx = T(1) / x;
    if (n == -2) return rx * rx;
    if (m_abs(x) <= T(3.4028234e38)) xnlo = -m_fma(x, rx, T(-1)) * rx;
    x = rx;
    n = -n;
  }
  unsigned int m = (unsigned int)n;
  while (m > 1u) {
    if (m & 1u) {
      T err = m_fma(y, xnlo, x * ynlo);
      two_mul(x, y, &y, &ynlo);
      ynlo += err;
    }
    T err = x * T(2) * xnlo;
    two_mul(x, x, &x, &xnlo);
    xnlo += err;
    m >>= 1;
  }
  T err = m_fma(y, xnlo, x * ynlo);
  T r = m_fma(x, y, err);
  bool fin = (m_abs(x) <= T(3.4028234e38)) && (m_abs(err) <= T(3.4028234e38));
  return fin ? r : x * y;
}

// ---- op tables ---------------------------------------------------------------
template <int OP, typename T> MCM_DI T unary_apply(T x) {
  if constexpr (OP == MCM_U_PLUS || OP == MCM_U_FLOAT) return x;
  else if constexpr (OP == MCM_U_NEG) return -x;
  else if constexpr (OP == MCM_U_EXP) return m_exp(x);
  else if constexpr (OP == MCM_U_EXP2) return m_exp2(x);
  else if constexpr (OP == MCM_U_EXP10) return m_exp10(x);
  else if constexpr (OP == MCM_U_EXPM1) return m_expm1(x);
  else if constexpr (OP == MCM_U_LOG) return m_log(x);
  else if constexpr (OP == MCM_U_LOG10) return m_log10(x);
  else if constexpr (OP == MCM_U_LOG2) return m_log2(x);
  else if constexpr (OP == MCM_U_LOG1P) return m_log1p(x);
  else if constexpr (OP == MCM_U_SIN) return m_sin(x);
  else if constexpr (OP == MCM_U_COS) return m_cos(x);
  else if constexpr (OP == MCM_U_TAN) return m_tan(x);
  else if constexpr (OP == MCM_U_SIND) return m_sind(x);
  else if constexpr (OP == MCM_U_COSD) return m_cosd(x);
  else if constexpr (OP == MCM_U_TAND) return m_tand(x);
  else if constexpr (OP == MCM_U_SINH) return m_sinh(x);
  else if constexpr (OP == MCM_U_COSH) return m_cosh(x);
  else if constexpr (OP == MCM_U_TANH) return m_tanh(x);
  else if constexpr (OP == MCM_U_ASIN) return m_asin(x);
  else if constexpr (OP == MCM_U_ACOS) return m_acos(x);
  else if constexpr (OP == MCM_U_ATAN) return m_atan(x);
  else if constexpr (OP == MCM_U_ASIND) return m_asin(x) * C<T>::rad2deg;
  else if constexpr (OP == MCM_U_ACOSD) return m_acos(x) * C<T>::rad2deg;
  else if constexpr (OP == MCM_U_ATAND) return m_atan(x) * C<T>::rad2deg;
  else if constexpr (OP == MCM_U_ASINH) return m_asinh(x);
  else if constexpr (OP == MCM_U_ACOSH) return m_acosh(x);
  else if constexpr (OP == MCM_U_ATANH) return m_atanh(x);
  else if constexpr (OP == MCM_U_ZERO) return T(0);
  else if constexpr (OP == MCM_U_SIGN) return m_sign(x);
  else if constexpr (OP == MCM_U_ABS) return m_abs(x);
  else if constexpr (OP == MCM_U_SQRT) return m_sqrt(x);
  else if constexpr (OP == MCM_U_RAD2DEG) return x * C<T>::rad2deg;
  else if constexpr (OP == MCM_U_DEG2RAD) return x * C<T>::deg2rad;
  else if constexpr (OP == MCM_U_ROUND) return m_rint(x);
  else if constexpr (OP == MCM_U_FLOOR) return m_floor(x);
  else if constexpr (OP == MCM_U_CEIL) return m_ceil(x);
  else if constexpr (OP == MCM_U_TRUNC) return m_trunc(x);
  else if constexpr (OP == MCM_U_INV) return T(1) / x;
  else if constexpr (OP == MCM_U_SQR) return x * x;
  else if constexpr (OP == MCM_U_CUBE) return x * x * x;
  else if constexpr (OP == MCM_U_ROUND_TIESAWAY) return m_round_tiesaway(x);
  else if constexpr (OP == MCM_U_ROUND_TIESUP) return m_round_tiesup(x);
  else if constexpr (OP == MCM_U_ROUND_FROMZERO) return m_round_fromzero(x);
  else if constexpr (OP == MCM_U_REM2PI_NEAREST) return m_rem2pi(x, 0);
  else if constexpr (OP == MCM_U_REM2PI_TOZERO) return m_rem2pi(x, 1);
  else if constexpr (OP == MCM_U_REM2PI_DOWN) return m_rem2pi(x, 2);
  else if constexpr (OP == MCM_U_REM2PI_UP) return m_rem2pi(x, 3);
  else return x;
}

// ops whose Base scalar version throws DomainError (H7): we flag NaN-from-non-NaN
template <int OP> struct UnaryCanDomainError {
  static constexpr bool value =
      OP == MCM_U_LOG || OP == MCM_U_LOG10 || OP == MCM_U_LOG2 || OP == MCM_U_LOG1P ||
      OP == MCM_U_SQRT || OP == MCM_U_ASIN || OP == MCM_U_ACOS || OP == MCM_U_ASIND ||
      OP == MCM_U_ACOSD || OP == MCM_U_ACOSH || OP == MCM_U_ATANH || OP == MCM_U_SIN ||
      OP == MCM_U_COS || OP == MCM_U_TAN || OP == MCM_U_SIND || OP == MCM_U_COSD ||
      OP == MCM_U_TAND || OP == MCM_U_REM2PI_NEAREST || OP == MCM_U_REM2PI_TOZERO || OP == MCM_U_REM2PI_DOWN ||
      OP == MCM_U_REM2PI_UP;
};

template <int OP, typename T> MCM_DI T binary_apply(T x, T y) {
  if constexpr (OP == MCM_B_ADD || OP == MCM_B_ADD_SUM) return x + y;
  else if constexpr (OP == MCM_B_SUB) return x - y;
  else if constexpr (OP == MCM_B_MUL) return x * y;
  else if constexpr (OP == MCM_B_DIV) return x / y;
  else if constexpr (OP == MCM_B_POW) return m_pow(x, y);
  else if constexpr (OP == MCM_B_MAX) return m_max(x, y);
  else if constexpr (OP == MCM_B_MIN) return m_min(x, y);
  else if constexpr (OP == MCM_B_MOD) return m_mod(x, y);
  else if constexpr (OP == MCM_B_MOD1) return m_mod1(x, y);
  else if constexpr (OP == MCM_B_ATAN2) return m_atan2(x, y);
  else if constexpr (OP == MCM_B_ATAND2) return m_atan2(x, y) * C<T>::rad2deg;
  else if constexpr (OP == MCM_B_HYPOT) return m_hypot(x, y);
  else if constexpr (OP == MCM_B_REM) return m_fmod(x, y);
  else if constexpr (OP == MCM_B_FDIV) return m_fdiv(x, y);
  else if constexpr (OP == MCM_B_REM_NEAREST) return m_remainder(x, y);
  else if constexpr (OP == MCM_B_REM_DOWN) return m_mod(x, y);
  else if constexpr (OP == MCM_B_REM_UP) return m_rem_up(x, y);
  else if constexpr (OP == MCM_B_REM_FROMZERO) return m_rem_fromzero(x, y);
  else if constexpr (OP == MCM_B_DIV_NEAREST) return m_div_r(x, y, m_remainder(x, y));
  else if constexpr (OP == MCM_B_DIV_DOWN) return m_div_r(x, y, m_mod(x, y));
  else if constexpr (OP == MCM_B_DIV_UP) return m_div_r(x, y, m_rem_up(x, y));
  else if constexpr (OP == MCM_B_DIV_FROMZERO) return m_div_r(x, y, m_rem_fromzero(x, y));
  else return x;
}
template <int OP> struct BinaryCanDomainError {
  static constexpr bool value = OP == MCM_B_POW || OP == MCM_B_MOD || OP == MCM_B_MOD1 ||
                                OP == MCM_B_REM || OP == MCM_B_FDIV || (OP >= MCM_B_REM_NEAREST && OP <= MCM_B_DIV_FROMZERO);
};

template <int OP, typename T> MCM_DI T ternary_apply(T x, T y, T z) {
  if constexpr (OP == MCM_T_CLAMP) return m_clamp(x, y, z);
  else if constexpr (OP == MCM_T_MULADD) return m_fma(x, y, z);
  else if constexpr (OP == MCM_T_MUL_ADD) return x * y + z;  // -fmad=false: two roundings
  else return x;
}

}  // namespace math
}  // namespace mcm
