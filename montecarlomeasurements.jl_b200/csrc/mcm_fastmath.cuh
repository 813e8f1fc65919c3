// mcm_fastmath.cuh — hand-written Float64 kernels for the hot transcendental
// primitives of the headline chain: sin, cos, exp, and `^` with a half-integer
// exponent.  They stand where the reference calls SLEEFPirates.sin_fast /
// cos_fast / exp (src/sleefpirates.jl:2-36) and Base `^` (src/particles.jl:183).
//
// Design target: <= 1 ulp from the correctly rounded result on the fast path
// (measured on the host twin against mpmath / long double, tests/test_fastmath_twin.py),
// with as few FP64-pipe instructions as possible — on B200 the fused Float64
// chain is bound by the FP64 pipe and issue slots, not by HBM.
// Anything outside a fast path's domain falls through to the CUDA math library.
//
// The file also compiles as plain C++ (MCM_HOST_TWIN) so that every routine can be
// checked bit-for-bit on the CPU: all arithmetic is IEEE double with explicit
// fma(), the library is built with -fmad=false, so host and device agree exactly.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

// Polynomial coefficients live in __constant__ memory so that they are read as constant-bank
// operands of DFMA (no instruction); as literals every 64-bit constant costs two UMOV issue slots.
#ifdef MCM_HOST_TWIN
#define MCM_FM_CONST static const
#define MCM_FM_HD inline
namespace mcm { namespace fast {
inline int64_t d2bits(double x) { int64_t b; memcpy(&b, &x, 8); return b; }
inline double bits2d(int64_t b) { double x; memcpy(&x, &b, 8); return x; }
inline double lib_sin(double x) { return ::sin(x); }
inline double lib_cos(double x) { return ::cos(x); }
inline double lib_exp(double x) { return ::exp(x); }
inline double lib_pow(double x, double y) { return ::pow(x, y); }
}}
#else
#define MCM_FM_CONST static __constant__
#define MCM_FM_HD __device__ __forceinline__
namespace mcm { namespace fast {
__device__ __forceinline__ int64_t d2bits(double x) { return __double_as_longlong(x); }
__device__ __forceinline__ double bits2d(int64_t b) { return __longlong_as_double(b); }
// out-of-range arguments take the CUDA math library; kept out of line so the rare paths cost the hot
// loop neither registers nor instruction-cache footprint
static __device__ __noinline__ double lib_sin(double x) { return ::sin(x); }
static __device__ __noinline__ double lib_cos(double x) { return ::cos(x); }
static __device__ __noinline__ double lib_exp(double x) { return ::exp(x); }
static __device__ __noinline__ double lib_pow(double x, double y) { return ::pow(x, y); }
}}
#endif

namespace mcm {
namespace fast {

// ---------------------------------------------------------------------------------
// sin / cos
// ---------------------------------------------------------------------------------
// Argument reduction x = n*(pi/2) + (r + rl), |r| <= pi/4 (+ rounding), valid for |x| <= 2^17*pi/2:
// pi/2 = P1 + P2 + P3 with P1 holding 33 bits, so n*P1 (n < 2^18) and r0 = x - n*P1 are exact.
// n*P2 is formed exactly as th + tl (one FMA); r1 = r0 - th is exact whenever the cancellation is
// heavy (Sterbenz), i.e. exactly in the cases where x sits next to a multiple of pi/2, and the
// Fast2Sum remainder picks up its rounding error otherwise; P3 supplies the tail.  Worst-case
// doubles near multiples of pi/2 therefore keep full relative accuracy — no slow path is needed.
MCM_FM_CONST double kTrig[20] = {
    6.36619772367581382433e-01,   // [0] 2/pi
    1.57079632673412561417e+00,   // [1] P1: first 33 bits of pi/2
    6.07710050650619224932e-11,   // [2] P2: pi/2 - P1, rounded to 53 bits
    3.52155986518320018165e-27,   // [3] P3: pi/2 - P1 - P2 (scripts/fit_fastmath.py)
    // sin kernel S6..S1 (fdlibm __kernel_sin)
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, 1.66666666666666324348e-01,
    // cos kernel C6..C1 (fdlibm __kernel_cos)
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
    205887.0,                      // [16] fast-path limit 2^17 * pi/2
    -1.66666666666666324348e-01,   // [17] S1 with its sign (sincos_kernel)
    0.0, 0.0};
#define kTwoOverPi kTrig[0]
#define kPio2_1 kTrig[1]
#define kPio2_2 kTrig[2]
#define kPio2_3 kTrig[3]
#define kTrigFastMax kTrig[16]

struct Reduced { double r, rl; int q; };
MCM_FM_HD bool is_zero_bits(double x) { return (d2bits(x) << 1) == 0; }

MCM_FM_HD Reduced trig_reduce_full(double x) {
  Reduced o;
  const double n = rint(x * kTwoOverPi);
  const double r0 = fma(-n, kPio2_1, x);          // exact
  const double th = n * kPio2_2;
  const double tl = fma(n, kPio2_2, -th);         // n*P2 == th + tl exactly
  const double r1 = r0 - th;
  const double d = (r0 - r1) - th;                // rounding error of r1 (Fast2Sum)
  o.r = r1;
  o.rl = fma(-n, kPio2_3, d - tl);
  o.q = (int)n;
  return o;
}
// Four FP64 instructions cheaper: r = RN(r0 - n P2) by one FMA and its rounding error by a second one.  The second
// FMA is exact only while r0 - r is representable, i.e. while r has not cancelled against n P2 (|n P2| <= 2^-16):
// for |r| >= 2^-11 the two are within a factor of two (Sterbenz).  Arguments that close to a NONZERO multiple of pi/2
// — one in ~3000 of uniformly spread ones — take trig_reduce_full (`ok` = false).  n == 0 (|x| < pi/4: small angles,
// the common case of a pendulum near rest) never cancels: r = x exactly, whatever its size.
MCM_FM_HD Reduced trig_reduce_cheap(double x, bool& ok) {
  Reduced o;
  const double n = rint(x * kTwoOverPi);
  const double r0 = fma(-n, kPio2_1, x);          // exact
  const double r = fma(-n, kPio2_2, r0);
  const double e = fma(-n, kPio2_2, r0 - r);      // rounding error of r (exact when ok)
  o.r = r;
  o.rl = fma(-n, kPio2_3, e);
  o.q = (int)n;
  ok = (d2bits(r) & 0x7fffffffffffffffLL) >= 0x3f40000000000000LL || o.q == 0;   // |r| >= 2^-11, or no multiple removed
  return o;
}
MCM_FM_HD Reduced trig_reduce(double x) {
  bool ok;
  const Reduced a = trig_reduce_cheap(x, ok);
  return ok ? a : trig_reduce_full(x);
}
// (-1)^(q>>1) applied on the integer pipe
MCM_FM_HD double flip_sign_if(double v, int bit2) { return bits2d(d2bits(v) ^ ((int64_t)(bit2 & 2) << 62)); }

// minimax kernels on |r| <= pi/4 (coefficients of the classic fdlibm __kernel_sin/__kernel_cos)
MCM_FM_HD double sin_kernel(double r, double rl) {
  const double z = r * r;
  const double v = z * r;
  double p = kTrig[4];
  p = fma(p, z, kTrig[5]);
  p = fma(p, z, kTrig[6]);
  p = fma(p, z, kTrig[7]);
  p = fma(p, z, kTrig[8]);
  // r - ((z*(rl/2 - v*p) - rl) - v*S1)
  const double t = fma(z, fma(-v, p, 0.5 * rl), -rl);
  return r - fma(v, kTrig[9], t);
}
MCM_FM_HD double cos_kernel(double r, double rl) {
  const double z = r * r;
  double p = kTrig[10];
  p = fma(p, z, kTrig[11]);
  p = fma(p, z, kTrig[12]);
  p = fma(p, z, kTrig[13]);
  p = fma(p, z, kTrig[14]);
  p = fma(p, z, kTrig[15]);
  // 1 - (z/2 - (z*z*p - r*rl)), arranged so that the two large terms cancel exactly
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + fma(z * z, p, -r * rl));
}

// sin(r + rl) (C == false) or cos(r + rl) (C == true) in ONE common form: both kernels above are
//   A + (v P(z) + B),  sin: A = r, v = z r, P = S1..S6, B = rl - (z/2) rl;  cos: A = w = 1 - z/2, v = z z, P = C1..C6,
//   B = ((1 - w) - z/2) - r rl
// (11 resp. 14 FP64 instructions, the leading terms r resp. 1 - z/2 exact or compensated: <= 0.80 ulp over the tested
// ranges).  sincos_kernel_t<C> fixes the parity at compile time; sincos_kernel selects the operands by the parity at
// run time — the SAME arithmetic on the selected operands, so all three return the same bits.  A warp whose lanes
// all need the same kernel (the usual case: an uncertain angle is narrow compared with pi/2) runs the fixed form;
// a warp with mixed parities runs the selecting form once (14 FP64 instructions + operand selects) instead of
// both kernels under divergence (25).
template <bool C>
MCM_FM_HD double sincos_kernel_t(double r, double rl) {
  const double z = r * r;
  const double hz = 0.5 * z;
  double p = C ? kTrig[10] : kTrig[4];
  p = fma(p, z, C ? kTrig[11] : kTrig[5]);
  p = fma(p, z, C ? kTrig[12] : kTrig[6]);
  p = fma(p, z, C ? kTrig[13] : kTrig[7]);
  p = fma(p, z, C ? kTrig[14] : kTrig[8]);
  p = fma(p, z, C ? kTrig[15] : kTrig[17]);
  if (C) {
    const double w = 1.0 - hz;
    const double e = (1.0 - w) - hz;
    return w + fma(z * z, p, fma(-r, rl, e));
  }
  return r + fma(z * r, p, fma(-hz, rl, rl));
}
MCM_FM_HD double sincos_kernel(double r, double rl, int odd) {
  const bool c = (odd & 1) != 0;
  const double z = r * r;
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  const double e = (1.0 - w) - hz;
  double p = c ? kTrig[10] : kTrig[4];
  p = fma(p, z, c ? kTrig[11] : kTrig[5]);
  p = fma(p, z, c ? kTrig[12] : kTrig[6]);
  p = fma(p, z, c ? kTrig[13] : kTrig[7]);
  p = fma(p, z, c ? kTrig[14] : kTrig[8]);
  p = fma(p, z, c ? kTrig[15] : kTrig[17]);
  const double v = z * (c ? z : r);
  const double b = fma(-(c ? r : hz), rl, c ? e : rl);
  return (c ? w : r) + fma(v, p, b);
}
MCM_FM_HD double sincos_branch(double r, double rl, int odd) {   // per-lane form of the exact paths
  return (odd & 1) ? sincos_kernel_t<true>(r, rl) : sincos_kernel_t<false>(r, rl);
}

MCM_FM_HD double sin_f64(double x) {
  if (!(fabs(x) <= kTrigFastMax)) return lib_sin(x);  // huge, Inf, NaN: library (Payne-Hanek)
  const Reduced a = trig_reduce(x);
  const double y = flip_sign_if(sincos_branch(a.r, a.rl, a.q), a.q);
  return is_zero_bits(x) ? x : y;                      // keep -0.0 (integer test: the FP64 pipe is the scarce one)
}
MCM_FM_HD double cos_f64(double x) {
  if (!(fabs(x) <= kTrigFastMax)) return lib_cos(x);
  const Reduced a = trig_reduce(x);
  return flip_sign_if(sincos_branch(a.r, a.rl, a.q + 1), a.q + 1);
}

// ---------------------------------------------------------------------------------
// exp
// ---------------------------------------------------------------------------------
// Table-driven (round 2): x = k (ln2/32) + r with k = rint(x 32/ln2), |r| <= ln2/64 (two-constant Cody-Waite, L1 holds 36
// bits so k L1 is exact, the rounding error of the second step is recovered into rl); k = 32 m + j;
//   exp(x) = 2^m * T_j * (1 + p),   T_j = 2^(j/32) = th + tl (mcm_exp_table.inc),   p = expm1(r + rl) = r + (r^2 q(r) + rl)
// with q the degree-4 Taylor polynomial of (expm1(r) - r)/r^2 (truncation r^7/5040 < 2^-57 on this range); the result is
// th + (tl + th p): <= 0.55 ulp measured on the host twin (0.83 before).  14 FP64 instructions against the 19 of the degree-10 polynomial on |r| <= ln2/2 it
// replaces (the FP64 pipe is what bounds the fused chains), plus one 16-byte table load.  |x| > 708: library.
#include "mcm_fastmath_coeffs.inc"  // MCM_EXPQ_LIST: kExpQ0 .. kExpQ10 (the former polynomial; kept for exp10 / tests)
#include "mcm_exp_table.inc"
MCM_FM_CONST double kExp[16] = {
    MCM_EXP_INV_L,                // [0] 32 / ln2
    MCM_EXP_L1,                   // [1] ln2/32, leading 36 bits
    MCM_EXP_L2,                   // [2] ln2/32 - L1
    708.0,                        // [3] fast-path limit
    5.00000000000000000000e-01, 1.66666666666666657415e-01, 4.16666666666666643537e-02, 8.33333333333333321769e-03,
    1.38888888888888894189e-03,   // [4..8] 1/2!, 1/3!, 1/4!, 1/5!, 1/6!
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#define kExpInvL kExp[0]
#define kExpL1 kExp[1]
#define kExpL2 kExp[2]
// The table is read with a per-lane index: __constant__ memory would serialise the 32 lanes of a warp, so on the device
// it lives in global memory and is read through L1 (one 16-byte load per value, 512 bytes in all).
#ifdef MCM_HOST_TWIN
static const double kExpTab[64] = {MCM_EXP_TABLE};
inline void exp_tab(int j, double& th, double& tl) { th = kExpTab[2 * j]; tl = kExpTab[2 * j + 1]; }
#else
static __device__ const double2 kExpTab[32] = {
#define MCM_PAIR_(a, b) {a, b}
    MCM_EXP_TABLE_PAIRS
#undef MCM_PAIR_
};
__device__ __forceinline__ void exp_tab(int j, double& th, double& tl) {
  const double2 t = __ldg(&kExpTab[j]);
  th = t.x; tl = t.y;
}
#endif

// the branch-free core, valid for |x| <= 708 (the callers test the range)
MCM_FM_HD double exp_core(double x) {
  const double k = rint(x * kExpInvL);
  const double rh = fma(-k, kExpL1, x);            // exact (k L1 has <= 53 bits)
  const double r = fma(-k, kExpL2, rh);
  const double rl = fma(-k, kExpL2, rh - r);       // rounding error of the line above
  const int ki = (int)k;
  double th, tl;
  exp_tab(ki & 31, th, tl);
  double q = kExp[8];
  q = fma(q, r, kExp[7]); q = fma(q, r, kExp[6]); q = fma(q, r, kExp[5]); q = fma(q, r, kExp[4]);
  const double p = r + fma(r * r, q, rl);          // expm1(r + rl)
  const double y = th + fma(th, p, tl);            // in (0.98, 2.02)
  return bits2d(d2bits(y) + ((int64_t)(ki >> 5) << 52));   // |m| <= 1022 and 2^m y stays normal for |x| <= 708
}
MCM_FM_HD double exp_f64(double x) {
  if (!(fabs(x) <= kExp[3])) return lib_exp(x);
  return exp_core(x);
}

// ---------------------------------------------------------------------------------
// x^y for half-integer y  (a^1.5 of the headline chain, a^0.5, a^2.5, a^-0.5 ...)
// ---------------------------------------------------------------------------------
// x^(m + 1/2) = x^m * sqrt(x).  sqrt is correctly rounded and its residual x - s*s is exactly
// representable, so sqrt(x) = s + rs with rs = (x - s^2)/(2s) known to ~2^-106; x^m is carried as an
// unevaluated pair by compensated power-by-squaring; one double-double product, one final rounding.
// Error < 1 ulp (typically 0.5).  General exponents use the CUDA library pow.
MCM_FM_HD void two_prod(double a, double b, double& hi, double& lo) { hi = a * b; lo = fma(a, b, -hi); }

MCM_FM_HD void powi_dd(double x, int m, double& hi, double& lo) {  // m >= 0
  double yh = 1.0, yl = 0.0, xh = x, xl = 0.0;
  while (m > 0) {
    if (m & 1) {
      double ph, pl;
      two_prod(yh, xh, ph, pl);
      yl = fma(yh, xl, fma(yl, xh, pl));
      yh = ph;
    }
    m >>= 1;
    if (m) {
      double sh, sl;
      two_prod(xh, xh, sh, sl);
      xl = fma(2.0 * xh, xl, sl);
      xh = sh;
    }
  }
  hi = yh; lo = yl;
}

MCM_FM_HD double pow_half_int(double x, int k) {  // k odd, exponent k/2; pow_half_window(x, k)
  const double s = sqrt(x);
  const double res = fma(-s, s, x);                 // exact
  const double rs = res * (0.5 / s);                // sqrt(x) = s + rs
  int m = (k - 1) / 2;                              // floor for k > 0
  if (k < 0) m = (-k + 1) / 2;                      // x^(k/2) = 1 / (x^m' / sqrt(x)) handled below
  double ph, pl;
  powi_dd(x, m, ph, pl);
  if (k > 0) {
    double hi, lo;
    two_prod(ph, s, hi, lo);
    lo = fma(ph, rs, fma(pl, s, lo));
    return hi + lo;
  }
  // negative: x^(-m' + 1/2) = sqrt(x) / x^m'
  const double q = s / ph;
  // one Newton correction in double-double: q += ((s + rs) - q*(ph + pl)) / ph
  const double rem = fma(-q, ph, s) + fma(-q, pl, rs);
  return q + rem / ph;
}

// x^1.5 = x*sqrt(x): p = x*s with its exact rounding error, plus x*rs ~= s*res/2 (x/s = s to 2^-53).
MCM_FM_HD double pow_three_halves(double x) {
  const double s = sqrt(x);
  const double res = fma(-s, s, x);
  const double p = x * s;
  const double pe = fma(x, s, -p);
  return p + fma(0.5 * res, s, pe);
}

// x > 0, normal, and neither x^(k/2) nor an intermediate power (x^((|k|+1)/2)) leaves the normal range:
// (|exponent(x)| + 1) (|k| + 1) < 2000.  Zero, denormals, Inf and NaN fail (|exponent| >= 1023).  Outside the
// window the compensated products overflow to Inf - Inf = NaN where the true power is Inf (or lose bits to
// denormals), so those arguments go to the library.
MCM_FM_HD bool pow_half_window(double x, int k) {
  const int e = (int)((d2bits(x) >> 52) & 0x7ff) - 1023;
  const int ae = e < 0 ? -e : e, ak = k < 0 ? -k : k;
  return x > 0.0 && (ae + 1) * (ak + 1) < 2000;
}
MCM_FM_HD double pow_f64(double x, double y) {
  const double t = y + y;
  if (fabs(t) <= 17.0 && t == rint(t)) {
    const int k = (int)t;
    if ((k & 1) && pow_half_window(x, k)) {
      if (k == 3) return pow_three_halves(x);
      if (k == 1) return sqrt(x);
      return pow_half_int(x, k);
    }
  }
  return lib_pow(x, y);
}

MCM_FM_HD double log_f64(double x) { return ::log(x); }

// ---------------------------------------------------------------------------------
// rem2pi / mod2pi  (Base.rem2pi(x, r::RoundingMode), src/particles.jl:359; mod2pi = rem2pi(x, RoundDown))
// ---------------------------------------------------------------------------------
// Base reduces x = n (pi/2) + y with y a double-double (rem_pio2_kernel: Cody-Waite for moderate x, Payne-Hanek
// beyond) and then adds the double-double multiple of pi/2 the rounding mode asks for, returning the high word of
// the sum (add22condh).  The same structure here: trig_reduce above for |x| <= 2^17 pi/2, an integer Payne-Hanek
// below for everything larger (1152 bits of 2/pi, 128 fraction bits kept), the mode logic restated from
// base/math.jl.  The result is the correctly rounded remainder except in near-tie cases (measured on the host twin
// against mpmath at 1500 bits: <= 0.5000001 ulp).
MCM_FM_CONST uint64_t kTwoOverPiBits[20] = {
    0xa2f9836e4e441529ULL, 0xfc2757d1f534ddc0ULL, 0xdb6295993c439041ULL, 0xfe5163abdebbc561ULL,
    0xb7246e3a424dd2e0ULL, 0x06492eea09d1921cULL, 0xfe1deb1cb129a73eULL, 0xe88235f52ebb4484ULL,
    0xe99c7026b45f7e41ULL, 0x3991d639835339f4ULL, 0x9c845f8bbdf9283bULL, 0x1ff897ffde05980fULL,
    0xef2f118b5a0a6d1fULL, 0x6d367ecf27cb09b7ULL, 0x4f463f669e5fea2dULL, 0x7527bac7ebe5f17bULL,
    0x3d0739f78a5292eaULL, 0x6bfb5fb11f8d5d08ULL, 0x56033046fc7b6babULL, 0xf0cfbc209af4361dULL};

MCM_FM_HD void umul128(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef MCM_HOST_TWIN
  const unsigned __int128 p = (unsigned __int128)a * b;
  hi = (uint64_t)(p >> 64); lo = (uint64_t)p;
#else
  lo = a * b; hi = __umul64hi(a, b);
#endif
}
MCM_FM_HD int clz64(uint64_t v) {
#ifdef MCM_HOST_TWIN
  return __builtin_clzll(v);
#else
  return __clzll((long long)v);
#endif
}
MCM_FM_HD uint64_t limb_window(const uint64_t (&q)[6], int pos) {   // bits [pos, pos + 64) of the 320-bit integer q
  const int limb = pos >> 6, off = pos & 63;
  uint64_t v = q[limb] >> off;
  if (off) v |= q[limb + 1] << (64 - off);
  return v;
}

// x = n (pi/2) + (yh + yl), |y| <= pi/4, for finite |x| >= 1; returns n mod 4.  Payne-Hanek on integers.
MCM_FM_HD int rem_pio2_large(double x, double& yh, double& yl) {
  const int64_t bits = d2bits(x);
  const int ex = (int)((bits >> 52) & 0x7ff);
  const uint64_t M = ((uint64_t)bits & 0x000fffffffffffffULL) | 0x0010000000000000ULL;
  const int E = ex - 1075;                              // |x| = M 2^E, -52 <= E <= 971
  int k0 = (E - 2) >> 6;                                // first 64-bit word of 2/pi that can matter mod 4
  if (k0 < 0) k0 = 0;
  uint64_t q[6] = {0, 0, 0, 0, 0, 0};                   // Q = M * (words k0 .. k0+3), little endian
  uint64_t carry = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint64_t hi, lo;
    umul128(M, kTwoOverPiBits[k0 + 3 - i], hi, lo);
    const uint64_t s = lo + carry;
    q[i] = s;
    carry = hi + (s < lo ? 1u : 0u);
  }
  q[4] = carry;
  const int b0 = 64 * k0 + 256 - E;                     // position of the units bit of |x| 2/pi in Q (191..308)
  int n = (int)(limb_window(q, b0) & 3u);
  uint64_t fh = limb_window(q, b0 - 64), fl = limb_window(q, b0 - 128);   // 128 fraction bits
  bool neg = false;
  if (fh >> 63) {                                       // fraction >= 1/2: round n up, y negative
    n = (n + 1) & 3;
    fl = ~fl + 1u; fh = ~fh + (fl == 0 ? 1u : 0u);
    neg = true;
  }
  int lz = 0;
  if (fh == 0) { fh = fl; fl = 0; lz = 64; }
  if (fh == 0) { yh = 0.0; yl = 0.0; return (x < 0.0) ? ((4 - n) & 3) : n; }
  const int z = clz64(fh);
  if (z) { fh = (fh << z) | (fl >> (64 - z)); fl <<= z; }
  lz += z;
  const double dh = (double)(fh >> 11);                                  // top 53 bits, exact
  const double dl = (double)(((fh & 0x7ffu) << 42) | (fl >> 22));        // next 53 bits, exact
  const double s1 = bits2d((int64_t)(1023 - 53 - lz) << 52);             // 2^(-53-lz)
  const double fhd = dh * s1, fld = (dl * s1) * 0x1p-53;                  // fraction = fhd + fld
  const double ph = 1.5707963267948966, pl = 6.123233995736766e-17;      // pi/2 as a double-double
  double h = fhd * ph;
  double l = fma(fhd, ph, -h) + fma(fhd, pl, fld * ph);
  double sh = h + l;
  l = l - (sh - h);
  if (neg != (x < 0.0)) { sh = -sh; l = -l; }
  yh = sh; yl = l;
  return (x < 0.0) ? ((4 - n) & 3) : n;
}

// n mod 4 and y = (yh, yl) with x = n pi/2 + y for any finite x
MCM_FM_HD int rem_pio2(double x, double& yh, double& yl) {
  if (fabs(x) <= kTrigFastMax) {
    const Reduced a = trig_reduce(x);
    const double h = a.r + a.rl;
    yl = a.rl - (h - a.r);
    yh = h;
    return a.q & 3;
  }
  return rem_pio2_large(x, yh, yl);
}

// high word of the sum of two double-doubles (Dekker; Base.Math.add22condh)
MCM_FM_HD double add22condh(double xh, double xl, double yh, double yl) {
  const double r = xh + yh;
  const double s = (fabs(xh) > fabs(yh)) ? (xh - r + yh + yl + xl) : (yh - r + xh + xl + yl);
  return r + s;
}
#define MCM_PI1O2_H 1.5707963267948966
#define MCM_PI1O2_L 6.123233995736766e-17
#define MCM_PI2O2_H 3.1415926535897931
#define MCM_PI2O2_L 1.2246467991473532e-16
#define MCM_PI3O2_H 4.7123889803846897
#define MCM_PI3O2_L 1.8369701987210297e-16
#define MCM_PI4O2_H 6.2831853071795862
#define MCM_PI4O2_L 2.4492935982947064e-16

// mode: 0 RoundNearest, 1 RoundToZero, 2 RoundDown (mod2pi), 3 RoundUp.  Inf -> NaN (DomainError in Base).
MCM_FM_HD double rem2pi_f64(double x, int mode) {
  if (!(fabs(x) <= 1.7976931348623157e308)) return x - x;
  double yh, yl;
  if (mode == 0) {
    if (fabs(x) < MCM_PI2O2_H) return x;
    const int n = rem_pio2(x, yh, yl);
    if ((n & 1) == 0) {
      if (n & 2) return yh <= 0.0 ? add22condh(yh, yl, MCM_PI2O2_H, MCM_PI2O2_L) : add22condh(yh, yl, -MCM_PI2O2_H, -MCM_PI2O2_L);
      return yh + yl;
    }
    return (n & 2) ? add22condh(yh, yl, -MCM_PI1O2_H, -MCM_PI1O2_L) : add22condh(yh, yl, MCM_PI1O2_H, MCM_PI1O2_L);
  }
  if (mode == 1) {
    const double ax = fabs(x);
    if (ax <= 6.283185307179585) return x;               // 2*Float64(pi, RoundDown)
    const int n = rem_pio2(ax, yh, yl);
    double z;
    if ((n & 1) == 0) {
      if (n & 2) z = add22condh(yh, yl, MCM_PI2O2_H, MCM_PI2O2_L);
      else z = yh > 0.0 ? yh + yl : add22condh(yh, yl, MCM_PI4O2_H, MCM_PI4O2_L);
    } else {
      z = (n & 2) ? add22condh(yh, yl, MCM_PI3O2_H, MCM_PI3O2_L) : add22condh(yh, yl, MCM_PI1O2_H, MCM_PI1O2_L);
    }
    return copysign(z, x);
  }
  if (mode == 2) {
    if (x < MCM_PI4O2_H) {
      if (x >= 0.0) return x;
      if (x > -MCM_PI4O2_H) return add22condh(x, 0.0, MCM_PI4O2_H, MCM_PI4O2_L);
    }
    const int n = rem_pio2(x, yh, yl);
    if ((n & 1) == 0) {
      if (n & 2) return add22condh(yh, yl, MCM_PI2O2_H, MCM_PI2O2_L);
      return yh > 0.0 ? yh + yl : add22condh(yh, yl, MCM_PI4O2_H, MCM_PI4O2_L);
    }
    return (n & 2) ? add22condh(yh, yl, MCM_PI3O2_H, MCM_PI3O2_L) : add22condh(yh, yl, MCM_PI1O2_H, MCM_PI1O2_L);
  }
  if (x > -MCM_PI4O2_H) {
    if (x <= 0.0) return x;
    if (x < MCM_PI4O2_H) return add22condh(x, 0.0, -MCM_PI4O2_H, -MCM_PI4O2_L);
  }
  const int n = rem_pio2(x, yh, yl);
  if ((n & 1) == 0) {
    if (n & 2) return add22condh(yh, yl, -MCM_PI2O2_H, -MCM_PI2O2_L);
    return yh < 0.0 ? yh + yl : add22condh(yh, yl, -MCM_PI4O2_H, -MCM_PI4O2_L);
  }
  return (n & 2) ? add22condh(yh, yl, -MCM_PI1O2_H, -MCM_PI1O2_L) : add22condh(yh, yl, -MCM_PI3O2_H, -MCM_PI3O2_L);
}


}  // namespace fast
}  // namespace mcm
