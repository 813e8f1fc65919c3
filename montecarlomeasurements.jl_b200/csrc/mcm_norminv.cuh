// mcm_norminv.cuh — quantile(Normal(0,1), p) on device, < 1 ulp.
//
// The reference evaluates Distributions.quantile -> StatsFuns.norminvcdf ->
// -erfcinv(2p)*sqrt(2) (SpecialFunctions' rational approximations), none of
// which is vendored; src/sampling.jl:10 calls it once per grid point.
//
// First version: Wichura's AS 241 (PPND16) evaluated in double — 5 ulp from the
// truth (two rationals of degree 7/7, a division, log and sqrt all round).  A
// Newton step cannot repair that in double: the residual Phi(z0) - p is needed to
// ~1e-17 relative.  This version evaluates plain polynomials whose LEADING terms are
// carried in double-double, so every rounding but the last is below 0.2 ulp:
//   centre |p - 1/2| <= 0.375:  z = q C(w), w = q^2, C = c0 + u T(u), u = w - wc
//   tails  pp = min(p, 1-p):    r = sqrt(-2 ln pp) as a double-double (ln by 2 atanh((m-1)/(m+1)) with
//                               a double-double leading term), piece k = exponent / top mantissa bit of
//                               r, s = r - centre_k,  |z| = g0 + g1 s + s^2 T_k(s) + G'(r) r_lo
// Coefficients: scripts/fit_norminv.py (Chebyshev-node interpolants at 70 digits, every piece within
// 2^-59 of the truth after rounding).  Measured on the host twin against mpmath: <= 0.75 ulp
// (tests/test_fastmath_twin.py); the device code is bit-identical to the twin (-fmad=false).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "mcm_norminv_coeffs.inc"

#ifdef MCM_HOST_TWIN
#define MCM_NI_HD inline
#define MCM_NI_TABLE static const
#define MCM_NI_CONST static const
namespace mcm { namespace norminv {
inline int ni_hi(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int ni_lo(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
inline double ni_make(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x; }
inline double ni_rcp32(double x) { return (double)(1.0f / (float)x); }
inline double ni_ld(const double* p) { return *p; }
}}
#else
#define MCM_NI_HD __device__ __forceinline__
#define MCM_NI_TABLE static __device__ const
#define MCM_NI_CONST static __constant__
namespace mcm { namespace norminv {
__device__ __forceinline__ int ni_hi(double x) { return __double2hiint(x); }
__device__ __forceinline__ int ni_lo(double x) { return __double2loint(x); }
__device__ __forceinline__ double ni_make(int hi, int lo) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ double ni_rcp32(double x) { return (double)__frcp_rn((float)x); }
// the tail table is indexed per lane (pieces differ inside a warp when the probabilities are permuted):
// global memory through the read-only path keeps divergent rows at L1 speed, a __constant__ table would
// serialise them
__device__ __forceinline__ double ni_ld(const double* p) { return __ldg(p); }
}}
#endif

namespace mcm {
namespace norminv {

MCM_NI_CONST double kCentre[MCM_NI_CDEG] = {MCM_NI_CENTRE_LIST};                 // c1 .. c_deg
MCM_NI_TABLE double kTail[MCM_NI_TPIECES][MCM_NI_TROW] = {MCM_NI_TAIL_ROWS};      // centre, g0h, g0l, g1h, g1l, c2..

// -2 ln(pp) as a double-double (th, tl), 0 < pp < 1/8 (pp may be subnormal).  Relative error ~2^-62.
MCM_NI_HD void neg2log_dd(double pp, double& th, double& tl) {
  int e = 0;
  if (pp < 2.2250738585072014e-308) { pp *= 18014398509481984.0; e = -54; }   // 2^54
  int hi = ni_hi(pp);
  e += (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;                 // m in [1, 2)
  if (hi >= 0x3ff6a09e) { hi -= 0x00100000; e += 1; }  // m in [sqrt(1/2), sqrt(2))
  const double m = ni_make(hi, ni_lo(pp));
  const double num = m - 1.0;                          // exact
  const double dh = m + 1.0;
  const double dl = m - (dh - 1.0);                    // exact rounding error of dh
  const double fh = num / dh;
  const double rr = fma(-fh, dh, num) - fh * dl;       // num - fh (dh + dl)
  const double fl = rr * ni_rcp32(dh);
  const double g = fh * fh;
  // 2 atanh(f) = 2 f + f g (2/3 + 2/5 g + 2/7 g^2 + ...), |f| <= 0.1716
  double s = 2.0 / 27.0;
  s = fma(s, g, 2.0 / 25.0); s = fma(s, g, 2.0 / 23.0); s = fma(s, g, 2.0 / 21.0); s = fma(s, g, 2.0 / 19.0);
  s = fma(s, g, 2.0 / 17.0); s = fma(s, g, 2.0 / 15.0); s = fma(s, g, 2.0 / 13.0); s = fma(s, g, 2.0 / 11.0);
  s = fma(s, g, 2.0 / 9.0); s = fma(s, g, 2.0 / 7.0); s = fma(s, g, 2.0 / 5.0); s = fma(s, g, 2.0 / 3.0);
  const double tail = fma(fh * g, s, 2.0 * fl);        // everything of ln m below 2 fh
  const double ed = (double)e;
  const double a = ed * 6.93147180369123816490e-01;    // e ln2_hi: exact (ln2_hi has 21 significant bits)
  const double b = 2.0 * fh;
  const double lh = a + b;                             // |a| >= 2.07 > |b|: Fast2Sum
  const double ll = (b - (lh - a)) + fma(ed, 1.90821492927058770002e-10, tail);
  const double h = lh + ll;
  const double l = ll - (h - lh);
  th = -2.0 * h; tl = -2.0 * l;
}

MCM_NI_HD double norminvcdf(double p) {
  if (!(p > 0.0 && p < 1.0)) {
    if (p == 0.0) return -__builtin_huge_val();
    if (p == 1.0) return __builtin_huge_val();
    return __builtin_nan("");
  }
  const double q = p - 0.5;
  if (fabs(q) <= 0.375) {
    const double ql = p - (q + 0.5);                   // p - 1/2 = q + ql exactly (Fast2Sum with |-1/2| >= p)
    const double w = fma(q + q, ql, q * q);
    const double u = w - MCM_NI_WC;
    double t = kCentre[MCM_NI_CDEG - 1];
#pragma unroll
    for (int k = MCM_NI_CDEG - 2; k >= 0; --k) t = fma(t, u, kCentre[k]);
    const double ch = fma(u, t, MCM_NI_C0H);
    const double cl = fma(u, t, MCM_NI_C0H - ch) + MCM_NI_C0L;   // C(w) = ch + cl
    return fma(q, ch, fma(q, cl, ql * ch));
  }
  const double pp = q < 0.0 ? p : 1.0 - p;             // exact for p > 1/2
  double th, tl;
  neg2log_dd(pp, th, tl);
  const double rh = sqrt(th);
  const double rl = (fma(-rh, rh, th) + tl) * (0.5 * ni_rcp32(rh));
  int k = (ni_hi(rh) >> 19) - 0x800;                   // [2,3) [3,4) [4,6) [6,8) ... [32,48)
  k = k < 0 ? 0 : (k > MCM_NI_TPIECES - 1 ? MCM_NI_TPIECES - 1 : k);
  const double* row = kTail[k];
  const double s = rh - ni_ld(row);                    // exact (Sterbenz)
  double t = ni_ld(row + MCM_NI_TROW - 1);
#pragma unroll
  for (int j = MCM_NI_TROW - 2; j >= 5; --j) t = fma(t, s, ni_ld(row + j));   // c2 + c3 s + ...
  const double g0h = ni_ld(row + 1), g0l = ni_ld(row + 2), g1h = ni_ld(row + 3), g1l = ni_ld(row + 4);
  const double ph = g1h * s;
  const double pl = fma(g1h, s, -ph) + g1l * s;
  const double zh = g0h + ph;                          // |g0| > |g1 s|: Fast2Sum
  const double ze = ph - (zh - g0h);
  const double slope = fma(2.0 * s, ni_ld(row + 5), g1h);        // G'(r) to ~1e-2
  const double zl = ((ze + pl) + g0l) + fma(s * s, t, slope * rl);
  const double z = zh + zl;
  return q < 0.0 ? -z : z;
}

}  // namespace norminv
}  // namespace mcm
