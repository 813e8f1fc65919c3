/*
 * mcm_oracle.c — CPU restatement of the MonteCarloMeasurements.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * montecarlomeasurements.jl_b200/ and libmcm_b200.so) may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker / the reported CPU baseline.
 *
 * The reference is pure Julia and cannot run here (no julia binary),
 * so this is a port, pinned against the reference's own known-answer values
 * (README.md:31-32, docs/src/index.md:55-62,73-74,346-354 — see
 * tests/test_oracle_golden.py).  Parity status per tier:
 *   exact tier  (+ - * / neg abs sqrt min max sign round rem mod clamp, literal
 *               powers, gather indices, order statistics, quantile formula):
 *               IEEE-754 semantics are language independent -> pinned by definition.
 *   ulp tier    (transcendentals, inverse normal CDF): "parity unpinned" — the
 *               reference calls SLEEFPirates / Base / SpecialFunctions, none of which
 *               is vendored or pinned (Project.toml:34-52 gives ranges only).  The
 *               oracle uses glibc libm (< 1 ulp) and an inverse normal CDF refined in
 *               long double; both are checked against mpmath in the tests.
 *   tolerance tier (mean/var/cov): long-double accumulations.
 *
 * All file:line citations are relative to /root/reference.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/mcm_b200.h"

#define ORC_API __attribute__((visibility("default")))

/* ---------- scalar semantics (Base Julia) ---------------------------------- */
static double j_sign(double x) { return x < 0 ? -1.0 : (x > 0 ? 1.0 : x); }
/* Base max/min for floats: NaN-propagating, -0.0 < +0.0 */
static double j_max(double x, double y) { double d = x - y; double a = signbit(d) ? y : x; return (isnan(x) || isnan(y)) ? d : a; }
static double j_min(double x, double y) { double d = x - y; double a = signbit(d) ? x : y; return (isnan(x) || isnan(y)) ? d : a; }
/* Base mod(x,y) for floats */
static double j_mod(double x, double y) {
  double r = fmod(x, y);
  if (r == 0) return copysign(r, y);
  if ((r > 0) != (y > 0)) return r + y;
  return r;
}
static double j_mod1(double x, double y) { double m = j_mod(x, y); return m == 0 ? y : m; }
static double j_fdiv(double x, double y) { return rint((x - fmod(x, y)) / y); }
static double j_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
/* round(x, r) for the modes beyond Nearest/ToZero/Up/Down (Base floatfuncs.jl); rem/div with a RoundingMode
 * (Base div.jl: rem(x,y,RoundDown) = mod(x,y), RoundUp = mod(x,-y), RoundNearest = IEEE remainder, RoundFromZero =
 * RoundUp when signs agree else RoundDown; div(x,y,r) = round((x - rem(x,y,r))/y)) — particles.jl:346,358,410-414 */
static double j_round_tiesaway(double x) { double y = trunc(x); return x == y ? y : trunc(2 * x - y); }
static double j_round_tiesup(double x) { const double e = 0x1p-53; return copysign(floor((x + (0.25 - e)) + (0.25 + e)), x); }
static double j_round_fromzero(double x) { return signbit(x) ? floor(x) : ceil(x); }
static double j_rem_up(double x, double y);
static double j_rem_fromzero(double x, double y);

static double j_rem_up(double x, double y) { return j_mod(x, -y); }
static double j_rem_fromzero(double x, double y) { return signbit(x) == signbit(y) ? j_rem_up(x, y) : j_mod(x, y); }
static float jf_sign(float x) { return x < 0 ? -1.0f : (x > 0 ? 1.0f : x); }
static float jf_max(float x, float y) { float d = x - y; float a = signbit(d) ? y : x; return (isnan(x) || isnan(y)) ? d : a; }
static float jf_min(float x, float y) { float d = x - y; float a = signbit(d) ? x : y; return (isnan(x) || isnan(y)) ? d : a; }
static float jf_mod(float x, float y) {
  float r = fmodf(x, y);
  if (r == 0) return copysignf(r, y);
  if ((r > 0) != (y > 0)) return r + y;
  return r;
}
static float jf_mod1(float x, float y) { float m = jf_mod(x, y); return m == 0 ? y : m; }
static float jf_fdiv(float x, float y) { return rintf((x - fmodf(x, y)) / y); }
static float jf_round_tiesaway(float x) { float y = truncf(x); return x == y ? y : truncf(2 * x - y); }
static float jf_round_tiesup(float x) { const float e = 0x1p-24f; return copysignf(floorf((x + (0.25f - e)) + (0.25f + e)), x); }
static float jf_round_fromzero(float x) { return signbit(x) ? floorf(x) : ceilf(x); }
static float jf_rem_up(float x, float y) { return jf_mod(x, -y); }
static float jf_rem_fromzero(float x, float y) { return signbit(x) == signbit(y) ? jf_rem_up(x, y) : jf_mod(x, y); }

/* sind/cosd: Base reduces in degrees exactly and converts with an extended pi/180;
 * long double evaluation of sin(x*pi/180) after an exact fmod is the arbiter here. */
static double j_sind(double x) {
  if (isinf(x) || isnan(x)) return NAN;
  double r = fmod(x, 360.0);
  if (r == 0) return r;
  if (fabs(r) == 180.0) return copysign(0.0, r);
  return (double)sinl((long double)r * 3.14159265358979323846264338327950288L / 180.0L);
}
static double j_cosd(double x) {
  if (isinf(x) || isnan(x)) return NAN;
  double r = fabs(fmod(x, 360.0));
  if (r == 90.0 || r == 270.0) return 0.0;
  return (double)cosl((long double)r * 3.14159265358979323846264338327950288L / 180.0L);
}

/* integer powers, src/particles.jl:383-391.  literal_pow(^,x,Val(k)) for k in
 * {0,1,2,3,-1,-2} expands to 1, x, x*x, x*x*x, inv(x), inv(x)^2 [upstream Base];
 * other k: long double power (the arbiter; Base's pow_body is < 1 ulp). */
static double j_powi(double x, int k) {
  switch (k) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return x * x;
    case 3: return x * x * x;
    case -1: return 1.0 / x;
    case -2: { double r = 1.0 / x; return r * r; }
    default: return (double)powl((long double)x, (long double)k);
  }
}
static float jf_powi(float x, int k) {
  switch (k) {
    case 0: return 1.0f;
    case 1: return x;
    case 2: return x * x;
    case 3: return x * x * x;
    case -1: return 1.0f / x;
    case -2: { float r = 1.0f / x; return r * r; }
    default: return (float)pow((double)x, (double)k);
  }
}

/* ---------- f(p): register_primitive_single, src/register_primitive.jl:94-97;
 *            list at src/particles.jl:185-190; SLEEF overrides src/sleefpirates.jl:2-36 */
/* Ulp-tier functions: the reference's own implementations (SLEEFPirates / Base, < 1 ulp class) are not vendored
 * (SURVEY.md H4), so the oracle returns the CORRECTLY ROUNDED value: x87 long double evaluation (64-bit mantissa)
 * rounded once to Float64; Float32 through Float64.  Device results are compared with these at <= 2 ulp. */
static double u64(int op, double x) {
  switch (op) {
    case MCM_U_PLUS: case MCM_U_FLOAT: return x;
    case MCM_U_NEG: return -x;
    case MCM_U_EXP: return (double)expl((long double)x);
    case MCM_U_EXP2: return (double)exp2l((long double)x);
    case MCM_U_EXP10: return (double)exp10l((long double)x);
    case MCM_U_EXPM1: return (double)expm1l((long double)x);
    case MCM_U_LOG: return (double)logl((long double)x);
    case MCM_U_LOG10: return (double)log10l((long double)x);
    case MCM_U_LOG2: return (double)log2l((long double)x);
    case MCM_U_LOG1P: return (double)log1pl((long double)x);
    case MCM_U_SIN: return (double)sinl((long double)x);
    case MCM_U_COS: return (double)cosl((long double)x);
    case MCM_U_TAN: return (double)tanl((long double)x);
    case MCM_U_SIND: return j_sind(x);
    case MCM_U_COSD: return j_cosd(x);
    case MCM_U_TAND: return j_sind(x) / j_cosd(x);
    case MCM_U_SINH: return (double)sinhl((long double)x);
    case MCM_U_COSH: return (double)coshl((long double)x);
    case MCM_U_TANH: return (double)tanhl((long double)x);
    case MCM_U_ASIN: return (double)asinl((long double)x);
    case MCM_U_ACOS: return (double)acosl((long double)x);
    case MCM_U_ATAN: return (double)atanl((long double)x);
    case MCM_U_ASIND: return (double)asinl((long double)x) * (180.0 / M_PI); /* rad2deg(asin(x)) */
    case MCM_U_ACOSD: return (double)acosl((long double)x) * (180.0 / M_PI);
    case MCM_U_ATAND: return (double)atanl((long double)x) * (180.0 / M_PI);
    case MCM_U_ASINH: return (double)asinhl((long double)x);
    case MCM_U_ACOSH: return (double)acoshl((long double)x);
    case MCM_U_ATANH: return (double)atanhl((long double)x);
    case MCM_U_ZERO: return 0.0;
    case MCM_U_SIGN: return j_sign(x);
    case MCM_U_ABS: return fabs(x);
    case MCM_U_SQRT: return sqrt(x);
    case MCM_U_RAD2DEG: return x * (180.0 / M_PI); /* Base: z * (180 / oftype(z, pi)) */
    case MCM_U_DEG2RAD: return x * (M_PI / 180.0); /* Base: z * (oftype(z, pi) / 180) */
    case MCM_U_ROUND: return rint(x);               /* RoundNearest, ties to even */
    case MCM_U_FLOOR: return floor(x);
    case MCM_U_CEIL: return ceil(x);
    case MCM_U_TRUNC: return trunc(x);
    case MCM_U_INV: return 1.0 / x;
    case MCM_U_SQR: return x * x;
    case MCM_U_CUBE: return x * x * x;
    case MCM_U_ROUND_TIESAWAY: return j_round_tiesaway(x);
    case MCM_U_ROUND_TIESUP: return j_round_tiesup(x);
    case MCM_U_ROUND_FROMZERO: return j_round_fromzero(x);
    default: return NAN;
  }
}
static float u32(int op, float x) {
  const float PI_F = 3.14159274f;
  switch (op) {
    case MCM_U_PLUS: case MCM_U_FLOAT: return x;
    case MCM_U_NEG: return -x;
    case MCM_U_EXP: return (float)exp((double)x);
    case MCM_U_EXP2: return (float)exp2((double)x);
    case MCM_U_EXP10: return (float)exp10((double)x);
    case MCM_U_EXPM1: return (float)expm1((double)x);
    case MCM_U_LOG: return (float)log((double)x);
    case MCM_U_LOG10: return (float)log10((double)x);
    case MCM_U_LOG2: return (float)log2((double)x);
    case MCM_U_LOG1P: return (float)log1p((double)x);
    case MCM_U_SIN: return (float)sin((double)x);
    case MCM_U_COS: return (float)cos((double)x);
    case MCM_U_TAN: return (float)tan((double)x);
    case MCM_U_SIND: return (float)j_sind((double)x);
    case MCM_U_COSD: return (float)j_cosd((double)x);
    case MCM_U_TAND: return (float)(j_sind((double)x) / j_cosd((double)x));
    case MCM_U_SINH: return (float)sinh((double)x);
    case MCM_U_COSH: return (float)cosh((double)x);
    case MCM_U_TANH: return (float)tanh((double)x);
    case MCM_U_ASIN: return (float)asin((double)x);
    case MCM_U_ACOS: return (float)acos((double)x);
    case MCM_U_ATAN: return (float)atan((double)x);
    case MCM_U_ASIND: return (float)asin((double)x) * (180.0f / PI_F);
    case MCM_U_ACOSD: return (float)acos((double)x) * (180.0f / PI_F);
    case MCM_U_ATAND: return (float)atan((double)x) * (180.0f / PI_F);
    case MCM_U_ASINH: return (float)asinh((double)x);
    case MCM_U_ACOSH: return (float)acosh((double)x);
    case MCM_U_ATANH: return (float)atanh((double)x);
    case MCM_U_ZERO: return 0.0f;
    case MCM_U_SIGN: return jf_sign(x);
    case MCM_U_ABS: return fabsf(x);
    case MCM_U_SQRT: return sqrtf(x);
    case MCM_U_RAD2DEG: return x * (180.0f / PI_F);
    case MCM_U_DEG2RAD: return x * (PI_F / 180.0f);
    case MCM_U_ROUND: return rintf(x);
    case MCM_U_FLOOR: return floorf(x);
    case MCM_U_CEIL: return ceilf(x);
    case MCM_U_TRUNC: return truncf(x);
    case MCM_U_INV: return 1.0f / x;
    case MCM_U_SQR: return x * x;
    case MCM_U_CUBE: return x * x * x;
    case MCM_U_ROUND_TIESAWAY: return jf_round_tiesaway(x);
    case MCM_U_ROUND_TIESUP: return jf_round_tiesup(x);
    case MCM_U_ROUND_FROMZERO: return jf_round_fromzero(x);
    default: return NAN;
  }
}

ORC_API int orc_unary(int op, int dtype, const void* in, void* out, size_t n) {
  if (op < 0 || op >= MCM_U_COUNT) return MCM_ERR_INVALID;
  if (dtype == MCM_F64) { const double* x = in; double* y = out; for (size_t i = 0; i < n; ++i) y[i] = u64(op, x[i]); }
  else if (dtype == MCM_F32) { const float* x = in; float* y = out; for (size_t i = 0; i < n; ++i) y[i] = u32(op, x[i]); }
  else return MCM_ERR_INVALID;
  return MCM_OK;
}

/* ---------- f(p1,p2), f(p,a), f(a,p): src/register_primitive.jl:44-58 ------- */
static double b64(int op, double x, double y) {
  switch (op) {
    case MCM_B_ADD: case MCM_B_ADD_SUM: return x + y;
    case MCM_B_SUB: return x - y;
    case MCM_B_MUL: return x * y;
    case MCM_B_DIV: return x / y;
    case MCM_B_POW: return (double)powl((long double)x, (long double)y);
    case MCM_B_MAX: return j_max(x, y);
    case MCM_B_MIN: return j_min(x, y);
    case MCM_B_MOD: return j_mod(x, y);
    case MCM_B_MOD1: return j_mod1(x, y);
    case MCM_B_ATAN2: return (double)atan2l((long double)x, (long double)y);
    case MCM_B_ATAND2: return (double)atan2l((long double)x, (long double)y) * (180.0 / M_PI);
    case MCM_B_HYPOT: return (double)hypotl((long double)x, (long double)y);
    case MCM_B_REM: return fmod(x, y);
    case MCM_B_FDIV: return j_fdiv(x, y);
    case MCM_B_REM_NEAREST: return remainder(x, y);
    case MCM_B_REM_DOWN: return j_mod(x, y);
    case MCM_B_REM_UP: return j_rem_up(x, y);
    case MCM_B_REM_FROMZERO: return j_rem_fromzero(x, y);
    case MCM_B_DIV_NEAREST: return rint((x - remainder(x, y)) / y);
    case MCM_B_DIV_DOWN: return rint((x - j_mod(x, y)) / y);
    case MCM_B_DIV_UP: return rint((x - j_rem_up(x, y)) / y);
    case MCM_B_DIV_FROMZERO: return rint((x - j_rem_fromzero(x, y)) / y);
    default: return NAN;
  }
}
static float b32(int op, float x, float y) {
  switch (op) {
    case MCM_B_ADD: case MCM_B_ADD_SUM: return x + y;
    case MCM_B_SUB: return x - y;
    case MCM_B_MUL: return x * y;
    case MCM_B_DIV: return x / y;
    case MCM_B_POW: return (float)pow((double)x, (double)y);
    case MCM_B_MAX: return jf_max(x, y);
    case MCM_B_MIN: return jf_min(x, y);
    case MCM_B_MOD: return jf_mod(x, y);
    case MCM_B_MOD1: return jf_mod1(x, y);
    case MCM_B_ATAN2: return (float)atan2((double)x, (double)y);
    case MCM_B_ATAND2: return atan2f(x, y) * (180.0f / 3.14159274f);
    case MCM_B_HYPOT: return (float)hypot((double)x, (double)y);
    case MCM_B_REM: return fmodf(x, y);
    case MCM_B_FDIV: return jf_fdiv(x, y);
    case MCM_B_REM_NEAREST: return remainderf(x, y);
    case MCM_B_REM_DOWN: return jf_mod(x, y);
    case MCM_B_REM_UP: return jf_rem_up(x, y);
    case MCM_B_REM_FROMZERO: return jf_rem_fromzero(x, y);
    case MCM_B_DIV_NEAREST: return rintf((x - remainderf(x, y)) / y);
    case MCM_B_DIV_DOWN: return rintf((x - jf_mod(x, y)) / y);
    case MCM_B_DIV_UP: return rintf((x - jf_rem_up(x, y)) / y);
    case MCM_B_DIV_FROMZERO: return rintf((x - jf_rem_fromzero(x, y)) / y);
    default: return NAN;
  }
}
ORC_API int orc_binary(int op, int dtype, const void* a, double sa, const void* b, double sb, void* out, size_t n) {
  if (op == MCM_B_RATDIV) return MCM_ERR_UNSUPPORTED;
  if (op < 0 || op >= MCM_B_COUNT) return MCM_ERR_INVALID;
  if (dtype == MCM_F64) {
    const double *x = a, *y = b; double* o = out;
    for (size_t i = 0; i < n; ++i) o[i] = b64(op, x ? x[i] : sa, y ? y[i] : sb);
  } else if (dtype == MCM_F32) {
    const float *x = a, *y = b; float* o = out;
    for (size_t i = 0; i < n; ++i) o[i] = b32(op, x ? x[i] : (float)sa, y ? y[i] : (float)sb);
  } else return MCM_ERR_INVALID;
  return MCM_OK;
}
/* clamp (particles.jl:183), muladd (particles.jl:416-422: fused allowed), x*y+z */
ORC_API int orc_ternary(int op, int dtype, const void* a, double sa, const void* b, double sb, const void* c, double sc,
                        void* out, size_t n) {
  if (dtype == MCM_F64) {
    const double *x = a, *y = b, *z = c; double* o = out;
    for (size_t i = 0; i < n; ++i) {
      double X = x ? x[i] : sa, Y = y ? y[i] : sb, Z = z ? z[i] : sc;
      if (op == MCM_T_CLAMP) o[i] = j_clamp(X, Y, Z);
      else if (op == MCM_T_MULADD) o[i] = fma(X, Y, Z);
      else if (op == MCM_T_MUL_ADD) { volatile double t = X * Y; o[i] = t + Z; }
      else return MCM_ERR_INVALID;
    }
  } else if (dtype == MCM_F32) {
    const float *x = a, *y = b, *z = c; float* o = out;
    for (size_t i = 0; i < n; ++i) {
      float X = x ? x[i] : (float)sa, Y = y ? y[i] : (float)sb, Z = z ? z[i] : (float)sc;
      if (op == MCM_T_CLAMP) o[i] = X > Z ? Z : (X < Y ? Y : X);
      else if (op == MCM_T_MULADD) o[i] = fmaf(X, Y, Z);
      else if (op == MCM_T_MUL_ADD) { volatile float t = X * Y; o[i] = t + Z; }
      else return MCM_ERR_INVALID;
    }
  } else return MCM_ERR_INVALID;
  return MCM_OK;
}
ORC_API int orc_powi(int dtype, const void* in, int k, void* out, size_t n) {
  if (dtype == MCM_F64) { const double* x = in; double* y = out; for (size_t i = 0; i < n; ++i) y[i] = j_powi(x[i], k); }
  else if (dtype == MCM_F32) { const float* x = in; float* y = out; for (size_t i = 0; i < n; ++i) y[i] = jf_powi(x[i], k); }
  else return MCM_ERR_INVALID;
  return MCM_OK;
}

/* ---------- statistics: src/particles.jl:302-323 ----------------------------- */
/* out = {n, mean, M2, min, max, sum}; long double two-pass (the arbiter for the
 * 1e-12 tolerance tier).  Statistics.mean is pairwise-sum/N, var is two-pass. */
ORC_API int orc_moments(int dtype, const void* in, size_t n, double* out) {
  /* shift by the first finite element so that the long double sums stay exact-ish even for
   * data like 1e15 +- 1e3 (a plain long double sum of 4e4 such values already exceeds 2^64) */
  long double s = 0.0L, sd = 0.0L;
  double mn = INFINITY, mx = -INFINITY, K = 0.0;
  int has_nan = 0;
  for (size_t i = 0; i < n; ++i) {
    double x = dtype == MCM_F64 ? ((const double*)in)[i] : (double)((const float*)in)[i];
    if (isfinite(x)) { K = x; break; }
  }
  for (size_t i = 0; i < n; ++i) {
    double x = dtype == MCM_F64 ? ((const double*)in)[i] : (double)((const float*)in)[i];
    s += x;
    sd += (long double)x - (long double)K;
    if (isnan(x)) has_nan = 1;
    if (x < mn) mn = x;
    if (x > mx) mx = x;
  }
  long double dm = n ? sd / (long double)n : NAN;
  long double m2 = 0.0L;
  for (size_t i = 0; i < n; ++i) {
    double x = dtype == MCM_F64 ? ((const double*)in)[i] : (double)((const float*)in)[i];
    long double d = ((long double)x - (long double)K) - dm;
    m2 += d * d;
  }
  out[0] = (double)n; out[1] = (double)((long double)K + dm); out[2] = (double)m2;
  out[3] = has_nan ? NAN : mn; out[4] = has_nan ? NAN : mx; out[5] = (double)s;
  return MCM_OK;
}
/* pvar/pstd: N == 1 -> zero(T) (particles.jl:308-312) */
ORC_API int orc_reduce(int kind, int dtype, const void* in, size_t n, double* out) {
  double m[6];
  if (n == 0) return MCM_ERR_INVALID;
  orc_moments(dtype, in, n, m);
  switch (kind) {
    case MCM_R_SUM: *out = m[5]; break;
    case MCM_R_MEAN: *out = m[1]; break;
    case MCM_R_VAR: *out = n == 1 ? 0.0 : m[2] / ((double)n - 1.0); break;
    case MCM_R_VARP: *out = n == 1 ? 0.0 : m[2] / (double)n; break;
    case MCM_R_STD: *out = n == 1 ? 0.0 : sqrt(m[2] / ((double)n - 1.0)); break;
    case MCM_R_STDP: *out = n == 1 ? 0.0 : sqrt(m[2] / (double)n); break;
    case MCM_R_MIN: *out = m[3]; break;
    case MCM_R_MAX: *out = m[4]; break;
    default: return MCM_ERR_INVALID;
  }
  return MCM_OK;
}
/* @prob p < a  (bymap.jl:128-136): mean(p.particles .< a) -> we return the count */
ORC_API int orc_count(int cmp, int dtype, const void* in, size_t n, double thr, uint64_t* out) {
  uint64_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    int b;
    if (dtype == MCM_F64) {
      double x = ((const double*)in)[i];
      b = cmp == MCM_CMP_LT ? x < thr : cmp == MCM_CMP_LE ? x <= thr : cmp == MCM_CMP_GT ? x > thr
          : cmp == MCM_CMP_GE ? x >= thr : cmp == MCM_CMP_EQ ? x == thr : x != thr;
    } else {
      float x = ((const float*)in)[i], t = (float)thr;
      b = cmp == MCM_CMP_LT ? x < t : cmp == MCM_CMP_LE ? x <= t : cmp == MCM_CMP_GT ? x > t
          : cmp == MCM_CMP_GE ? x >= t : cmp == MCM_CMP_EQ ? x == t : x != t;
    }
    c += b ? 1 : 0;
  }
  *out = c;
  return MCM_OK;
}

/* pcov(v) = cov(Matrix(v)) (particles.jl:476): centre columns, X'X / (N - corrected).
 * X given as M pointers to N-vectors (column j of Matrix(v) is variable j,
 * particles.jl:461-465).  Long double accumulation. */
ORC_API int orc_cov(int dtype, const void* const* vars, int M, size_t n, int corrected, double* C) {
  long double* mu = malloc(sizeof(long double) * (size_t)M);
  if (!mu) return MCM_ERR_NOMEM;
#define XV(j, i) (dtype == MCM_F64 ? ((const double*)vars[j])[i] : (double)((const float*)vars[j])[i])
  for (int j = 0; j < M; ++j) {
    long double s = 0;
    for (size_t i = 0; i < n; ++i) s += XV(j, i);
    mu[j] = s / (long double)n;
  }
  for (int a = 0; a < M; ++a)
    for (int b = a; b < M; ++b) {
      long double s = 0;
      for (size_t i = 0; i < n; ++i) s += ((long double)XV(a, i) - mu[a]) * ((long double)XV(b, i) - mu[b]);
      double v = (double)(s / ((long double)n - (corrected ? 1.0L : 0.0L)));
      C[(size_t)a * M + b] = v;
      C[(size_t)b * M + a] = v;
    }
#undef XV
  free(mu);
  return MCM_OK;
}

/* ---------- order statistics ---------------------------------------------------- */
/* Julia isless on floats: -0.0 < 0.0, NaN last */
static int isless64(double a, double b) {
  if (isnan(a)) return 0;
  if (isnan(b)) return 1;
  if (a == b) return signbit(a) && !signbit(b);
  return a < b;
}
static int cmp64(const void* pa, const void* pb) {
  double a = *(const double*)pa, b = *(const double*)pb;
  return isless64(a, b) ? -1 : (isless64(b, a) ? 1 : 0);
}
static int cmp32(const void* pa, const void* pb) {
  double a = *(const float*)pa, b = *(const float*)pb;
  return isless64(a, b) ? -1 : (isless64(b, a) ? 1 : 0);
}
ORC_API int orc_sort(int dtype, const void* in, void* out, size_t n) {
  if (dtype == MCM_F64) { memcpy(out, in, n * 8); qsort(out, n, 8, cmp64); }
  else if (dtype == MCM_F32) { memcpy(out, in, n * 4); qsort(out, n, 4, cmp32); }
  else return MCM_ERR_INVALID;
  return MCM_OK;
}
typedef struct { double v; int64_t i; } kv_t;
static int cmpkv(const void* pa, const void* pb) {
  const kv_t *a = pa, *b = pb;
  if (isless64(a->v, b->v)) return -1;
  if (isless64(b->v, a->v)) return 1;
  return a->i < b->i ? -1 : (a->i > b->i ? 1 : 0); /* stable */
}
ORC_API int orc_sortperm(int dtype, const void* in, int64_t* perm, size_t n) {
  kv_t* t = malloc(sizeof(kv_t) * (n ? n : 1));
  if (!t) return MCM_ERR_NOMEM;
  for (size_t i = 0; i < n; ++i) {
    t[i].v = dtype == MCM_F64 ? ((const double*)in)[i] : (double)((const float*)in)[i];
    t[i].i = (int64_t)i;
  }
  qsort(t, n, sizeof(kv_t), cmpkv);
  for (size_t i = 0; i < n; ++i) perm[i] = t[i].i;
  free(t);
  return MCM_OK;
}
/* pquantile(p,q) = Statistics.quantile(p.particles, q) (particles.jl:321), default
 * alpha = beta = 1 [upstream Statistics._quantile]:
 *   aleph = n*p + (1-p); j = clamp(trunc(aleph),1,n-1); g = clamp(aleph-j,0,1);
 *   a = v[j]; b = v[j+1]; a + g*(b-a)    (NaN present -> ArgumentError) */
ORC_API int orc_quantile(int dtype, const void* in, size_t n, const double* q, int nq, double* out) {
  if (n == 0) return MCM_ERR_INVALID;
  size_t esz = dtype == MCM_F64 ? 8 : 4;
  void* s = malloc(n * esz);
  if (!s) return MCM_ERR_NOMEM;
  orc_sort(dtype, in, s, n);
  double last = dtype == MCM_F64 ? ((double*)s)[n - 1] : (double)((float*)s)[n - 1];
  if (isnan(last)) { free(s); return MCM_ERR_INVALID; }
  for (int k = 0; k < nq; ++k) {
    double p = q[k];
    if (!(p >= 0.0 && p <= 1.0)) { free(s); return MCM_ERR_INVALID; }
    double aleph = (double)n * p + (1.0 - p);
    long long j = (long long)trunc(aleph);
    if (j > (long long)n - 1) j = (long long)n - 1;
    if (j < 1) j = 1;
    double g = aleph - (double)j;
    g = g < 0 ? 0 : (g > 1 ? 1 : g);
    if (dtype == MCM_F64) {
      double a = n == 1 ? ((double*)s)[0] : ((double*)s)[j - 1];
      double b = n == 1 ? ((double*)s)[0] : ((double*)s)[j];
      out[k] = (isfinite(a) && isfinite(b)) ? a + g * (b - a) : (1.0 - g) * a + g * b;
    } else {
      float a = n == 1 ? ((float*)s)[0] : ((float*)s)[j - 1];
      float b = n == 1 ? ((float*)s)[0] : ((float*)s)[j];
      volatile float d = b - a; /* Float32 subtraction, then promotion to Float64 */
      out[k] = (isfinite(a) && isfinite(b)) ? (double)a + g * (double)d : (1.0 - g) * (double)a + g * (double)b;
    }
  }
  free(s);
  return MCM_OK;
}

/* ---------- sampling: src/sampling.jl:6-13 ---------------------------------------- */
/* inverse normal CDF: AS241 start + two Newton steps on erfc in long double
 * (near correctly-rounded double; checked against mpmath in the tests). */
static long double as241l(long double p) {
  long double q = p - 0.5L, r, val;
  if (fabsl(q) <= 0.425L) {
    r = 0.180625L - q * q;
    return q * (((((((2.5090809287301226727e3L * r + 3.3430575583588128105e4L) * r + 6.7265770927008700853e4L) * r + 4.5921953931549871457e4L) * r + 1.3731693765509461125e4L) * r + 1.9715909503065514427e3L) * r + 1.3314166789178437745e2L) * r + 3.3871328727963666080e0L) /
           (((((((5.2264952788528545610e3L * r + 2.8729085735721942674e4L) * r + 3.9307895800092710610e4L) * r + 2.1213794301586595867e4L) * r + 5.3941960214247511077e3L) * r + 6.8718700749205790830e2L) * r + 4.2313330701600911252e1L) * r + 1.0L);
  }
  r = q < 0 ? p : 1.0L - p;
  r = sqrtl(-logl(r));
  if (r <= 5.0L) {
    r -= 1.6L;
    val = (((((((7.74545014278341407640e-4L * r + 2.27238449892691845833e-2L) * r + 2.41780725177450611770e-1L) * r + 1.27045825245236838258e0L) * r + 3.64784832476320460504e0L) * r + 5.76949722146069140550e0L) * r + 4.63033784615654529590e0L) * r + 1.42343711074968357734e0L) /
          (((((((1.05075007164441684324e-9L * r + 5.47593808499534494600e-4L) * r + 1.51986665636164571966e-2L) * r + 1.48103976427480074590e-1L) * r + 6.89767334985100004550e-1L) * r + 1.67638483018380384940e0L) * r + 2.05319162663775882187e0L) * r + 1.0L);
  } else {
    r -= 5.0L;
    val = (((((((2.01033439929228813265e-7L * r + 2.71155556874348757815e-5L) * r + 1.24266094738807843860e-3L) * r + 2.65321895265761230930e-2L) * r + 2.96560571828504891230e-1L) * r + 1.78482653991729133580e0L) * r + 5.46378491116411436990e0L) * r + 6.65790464350110377720e0L) /
          (((((((2.04426310338993978564e-15L * r + 1.42151175831644588870e-7L) * r + 1.84631831751005468180e-5L) * r + 7.86869131145613259100e-4L) * r + 1.48753612908506148525e-2L) * r + 1.36929880922735805310e-1L) * r + 5.99832206555887937690e-1L) * r + 1.0L);
  }
  return q < 0 ? -val : val;
}
ORC_API double orc_norminvcdf(double p) {
  if (!(p >= 0.0 && p <= 1.0)) return NAN;
  if (p == 0.0) return -INFINITY;
  if (p == 1.0) return INFINITY;
  /* work on the lower tail for relative accuracy: x = -Phi^-1(pt), pt = min(p, 1-p) */
  int upper = p > 0.5;
  long double pt = upper ? 1.0L - (long double)p : (long double)p;
  long double x = as241l(pt); /* <= 0 */
  for (int it = 0; it < 2; ++it) {
    long double cdf = 0.5L * erfcl(-x * 0.70710678118654752440084436210484903928L);
    long double pdf = 0.39894228040143267793994605993438186848L * expl(-0.5L * x * x);
    long double f = (cdf - pt) / pdf;
    x = x - f / (1.0L + 0.5L * x * f); /* Halley */
  }
  return (double)(upper ? -x : x);
}
/* systematic_sample(N, Normal(mu,sigma); permute): sampling.jl:8-11.  Grid point i
 * (0-based) y_i = (2i+1)/(2N) (H3: one correctly rounded division); quantile(Normal) =
 * mu + sigma*z.  perm (0-based, may be NULL) is the randperm of sampling.jl:11. */
ORC_API int orc_systematic_normal(int dtype, void* out, size_t n, double mu, double sigma, const int64_t* perm,
                                  uint64_t i0, uint64_t n_total) {
  for (size_t j = 0; j < n; ++j) {
    uint64_t i = perm ? (uint64_t)perm[j] : i0 + j;
    double p = (double)(2 * i + 1) / (2.0 * (double)n_total);
    double v = mu + sigma * orc_norminvcdf(p);
    if (dtype == MCM_F64) ((double*)out)[j] = v; else ((float*)out)[j] = (float)v;
  }
  return MCM_OK;
}
ORC_API int orc_systematic_uniform(int dtype, void* out, size_t n, double a, double b, const int64_t* perm,
                                   uint64_t i0, uint64_t n_total) {
  for (size_t j = 0; j < n; ++j) {
    uint64_t i = perm ? (uint64_t)perm[j] : i0 + j;
    double p = (double)(2 * i + 1) / (2.0 * (double)n_total);
    double v = a + (b - a) * p;
    if (dtype == MCM_F64) ((double*)out)[j] = v; else ((float*)out)[j] = (float)v;
  }
  return MCM_OK;
}

/* ---------- resampling: src/resampling.jl:76-84 ------------------------------------ */
ORC_API int orc_gather(int dtype, const void* src, const int64_t* idx, void* out, size_t n_draw, size_t n_src) {
  for (size_t j = 0; j < n_draw; ++j) {
    if (idx[j] < 0 || (size_t)idx[j] >= n_src) return MCM_ERR_INVALID; /* BoundsError */
    if (dtype == MCM_F64) ((double*)out)[j] = ((const double*)src)[idx[j]];
    else ((float*)out)[j] = ((const float*)src)[idx[j]];
  }
  return MCM_OK;
}
/* Philox-4x32-10 (the counter-based generator standing in for rand(rng,1:N), H5) */
static void philox(uint64_t ctr, uint64_t seed, uint32_t o[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
ORC_API int orc_rand_indices(int64_t* out, size_t n, uint64_t N, uint64_t seed, uint64_t offset) {
  for (size_t j = 0; j < n; ++j) {
    uint32_t r[4];
    philox(offset + j, seed, r);
    uint64_t u = ((uint64_t)r[1] << 32) | r[0];
    out[j] = (int64_t)(((unsigned __int128)u * N) >> 64);
  }
  return MCM_OK;
}
ORC_API int orc_randperm(int64_t* perm, size_t n, uint64_t seed) {
  kv_t* t = malloc(sizeof(kv_t) * (n ? n : 1));
  uint64_t* k = malloc(sizeof(uint64_t) * (n ? n : 1));
  if (!t || !k) return MCM_ERR_NOMEM;
  /* stable sort of 64-bit Philox keys */
  for (size_t j = 0; j < n; ++j) { uint32_t r[4]; philox(j, seed, r); k[j] = ((uint64_t)r[1] << 32) | r[0]; }
  /* sort indices by key (ties by index): simple LSD byte radix, stable */
  int64_t* a = malloc(sizeof(int64_t) * (n ? n : 1)); int64_t* b = malloc(sizeof(int64_t) * (n ? n : 1));
  for (size_t j = 0; j < n; ++j) a[j] = (int64_t)j;
  for (int pass = 0; pass < 8; ++pass) {
    size_t cnt[257] = {0};
    for (size_t j = 0; j < n; ++j) cnt[((k[a[j]] >> (8 * pass)) & 255) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (size_t j = 0; j < n; ++j) b[cnt[(k[a[j]] >> (8 * pass)) & 255]++] = a[j];
    int64_t* tmp = a; a = b; b = tmp;
  }
  memcpy(perm, a, n * sizeof(int64_t));
  free(a); free(b); free(t); free(k);
  return MCM_OK;
}
/* Systematic resampling — the dead reference routine src/resampling.jl:45-67:
 *   s = u*Sigma/N; for i: advance b while s >= bin (bin = running weight sum); x[i] = x[b]; s += Sigma/N
 * restated with the exact integer weights of include/mcm_b200.h (H10) so the indices are
 * independent of summation order: W = floor(w * 2^s), C = cumsum(W), S = C[n-1],
 * U = floor(u*2^32):  idx[i] = min{b : C_b * n * 2^32 > (i*2^32 + U) * S}. */
ORC_API int orc_resample_systematic(const double* w, size_t n, double u, int64_t* idx) {
  double wmax = 0;
  for (size_t i = 0; i < n; ++i) { if (!(w[i] >= 0)) return MCM_ERR_INVALID; if (w[i] > wmax) wmax = w[i]; }
  if (!(wmax > 0) || !isfinite(wmax)) return MCM_ERR_INVALID;
  int e; frexp(wmax, &e);
  int lg = 0; while (((uint64_t)1 << lg) < n) ++lg;
  double scale = ldexp(1.0, 62 - lg - e);
  uint64_t* C = malloc(sizeof(uint64_t) * n);
  uint64_t run = 0;
  for (size_t i = 0; i < n; ++i) { run += (uint64_t)floor(w[i] * scale); C[i] = run; }
  uint64_t S = run, U = (uint64_t)floor(u * 4294967296.0);
  size_t b = 0;
  for (size_t i = 0; i < n; ++i) {
    unsigned __int128 rhs = ((unsigned __int128)i * 4294967296ull + U) * S;
    while (!((unsigned __int128)C[b] * n * 4294967296ull > rhs)) ++b;
    idx[i] = (int64_t)b;
  }
  free(C);
  return MCM_OK;
}

/* ---- keyed index bijection (restates csrc/sampling.cu feistel_permute; the reference uses
 * randperm(rng, N) at src/sampling.jl:11, whose stream cannot be reproduced — H5).  Test
 * infrastructure: checks that the device evaluates exactly this permutation. ---- */
static uint64_t orc_splitmix64(uint64_t* s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint32_t orc_feistel_mix(uint32_t v, uint32_t k) {
  uint32_t h = v + k;
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}
ORC_API void orc_index_bijection(int64_t* out, size_t n, uint64_t seed, uint64_t column, uint64_t i0, uint64_t n_total) {
  uint32_t key[8];
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (column + 1));
  for (int r = 0; r < 8; ++r) key[r] = (uint32_t)(orc_splitmix64(&s) >> 32);
  uint32_t bits = 2;
  while (bits < 64 && (1ull << bits) < n_total) ++bits;
  const uint32_t lbits = bits / 2, rbits = bits - lbits;
  const uint32_t lmask = lbits >= 32 ? 0xffffffffu : ((1u << lbits) - 1u);
  const uint32_t rmask = rbits >= 32 ? 0xffffffffu : ((1u << rbits) - 1u);
  for (size_t j = 0; j < n; ++j) {
    uint64_t x = i0 + j;
    do {
      uint32_t L = (uint32_t)(x >> rbits) & lmask, R = (uint32_t)x & rmask;
      for (int r = 0; r < 4; r += 2) {   /* kFeistelRounds */
        L ^= orc_feistel_mix(R, key[r]) & lmask;
        R ^= orc_feistel_mix(L, key[r + 1]) & rmask;
      }
      x = ((uint64_t)L << rbits) | R;
    } while (x >= n_total);
    out[j] = (int64_t)x;
  }
}

/* ---- constant matrix x particle vector (src/particles.jl:688-707 `Matrix(b) * A'`,
 * src/sigmapoints.jl:71): y[j][i] = bias[j] + sum_k A[j,k] (x[k][i] - shift[k]) in long double.
 * xs: Mi x n row-major (variable-major), ys: Mo x n. ---- */
ORC_API void orc_pgemm(const double* A, int Mo, int Mi, const double* xs, const double* shift, const double* bias,
               double* ys, size_t n) {
  for (int j = 0; j < Mo; ++j)
    for (size_t i = 0; i < n; ++i) {
      long double acc = 0.0L;
      for (int k = 0; k < Mi; ++k)
        acc += (long double)A[(size_t)j * Mi + k] * ((long double)xs[(size_t)k * n + i] - (shift ? (long double)shift[k] : 0.0L));
      ys[(size_t)j * n + i] = (double)(acc + (bias ? (long double)bias[j] : 0.0L));
    }
}

/* ---- iid uniforms of mcm_rand_real (csrc/sampling.cu k_rand_real): u = (top 53 Philox bits + 1/2) 2^-53.
 * The reference draws rand(rng, d, N) from Julia's task-local generator (types.jl:71,79,93,97), whose stream
 * cannot be reproduced (H5); this restates the device stand-in for bit-exact tests. ---- */
ORC_API void orc_rand_uniform(double* out, size_t n, uint64_t seed, uint64_t offset) {
  for (size_t j = 0; j < n; ++j) {
    uint32_t r[4];
    philox(offset + j, seed, r);
    const uint64_t k = ((((uint64_t)r[1] << 32) | r[0]) >> 11);
    out[j] = ((double)k + 0.5) * 0x1p-53;
  }
}

/* ---- StatsBase.autocor(x, 1:maxlag; demean=true) [upstream StatsBase, called at src/sampling.jl:27]:
 * ac[k] = sum_{i<N-k} (x_i - m)(x_{i+k} - m) / sum_i (x_i - m)^2, long double accumulation. ---- */
ORC_API void orc_autocor(const double* x, size_t n, int maxlag, double* out) {
  long double m = 0.0L;
  for (size_t i = 0; i < n; ++i) m += x[i];
  m /= (long double)n;
  long double den = 0.0L;
  for (size_t i = 0; i < n; ++i) den += ((long double)x[i] - m) * ((long double)x[i] - m);
  for (int k = 1; k <= maxlag; ++k) {
    long double acc = 0.0L;
    for (size_t i = 0; i + (size_t)k < n; ++i) acc += ((long double)x[i] - m) * ((long double)x[i + k] - m);
    out[k - 1] = (double)(acc / den);
  }
}
