// mcm_dist.cuh — quantile(d, p) for the univariate distributions the sampling constructors accept.
//
// The reference builds `Particles(N, d)` for ANY continuous univariate Distribution by evaluating
// `quantile.((d,), y)` on the systematic grid (src/sampling.jl:6-13; constructors src/types.jl:67-90; `a ⊠ d`,
// `a ⊞ d` src/particles.jl:42,50).  Distributions.jl / StatsFuns / SpecialFunctions are not vendored,
// so these are the textbook inverse CDFs, stated per family below with Distributions' parameterisation:
//   closed forms      Normal, Uniform, Exponential, LogNormal, Cauchy, Laplace, Logistic, Rayleigh, Weibull, Pareto,
//                     Gumbel, TriangularDist, Frechet, Kumaraswamy
//   incomplete gamma  Gamma(alpha, theta), Chisq(nu), InverseGamma(alpha, theta), Erlang
//   incomplete beta   Beta(alpha, beta), TDist(nu), FDist(nu1, nu2)
// The incomplete-function inverses are a Halley iteration on the regularised function (series / modified-Lentz
// continued fraction), started from the classic closed-form guesses and solved on whichever of p, q = 1 - p is the
// smaller, so that both tails keep RELATIVE accuracy.  Callers pass p AND q: on the systematic grid both are exact
// rationals ((2i+1)/(2N) and (2(N-1-i)+1)/(2N)) and 1 - p would lose the upper tail.
// Tolerance tier: <= 1e-12 relative against scipy.special (tests/test_dist_twin.py on the host twin, which is the
// same source compiled as C++; tests/test_gpu_round2.py on the device).
#pragma once
#include <math.h>
#include <stdint.h>
#include "mcm_norminv.cuh"

#ifdef MCM_HOST_TWIN
#define MCM_DS_HD inline
#define MCM_DS_COLD inline
#else
#define MCM_DS_HD __device__ __forceinline__
#define MCM_DS_COLD static __device__ __noinline__     // the iterative inverses: one copy each, out of the switch
#endif

namespace mcm {
namespace dist {

enum Kind {
  NORMAL = 0, UNIFORM, EXPONENTIAL, LOGNORMAL, CAUCHY, LAPLACE, LOGISTIC, RAYLEIGH, WEIBULL, PARETO, GUMBEL, TRIANGULAR,
  FRECHET, KUMARASWAMY, GAMMA, CHISQ, INVERSEGAMMA, BETA, TDIST, FDIST, KIND_COUNT
};

struct Params { double a, b, c; };

// ---- regularised incomplete gamma P(a, x), Q(a, x) -------------------------------------------------------------
// prefactor x^a e^-x / Gamma(a + 1), evaluated in log space
MCM_DS_HD double gamma_prefactor(double a, double x) { return exp(a * log(x) - x - lgamma(a + 1.0)); }

MCM_DS_HD void gamma_pq(double a, double x, double& P, double& Q) {
  if (!(x > 0.0)) { P = 0.0; Q = 1.0; return; }
  if (x == __builtin_huge_val()) { P = 1.0; Q = 0.0; return; }
  const double pre = gamma_prefactor(a, x);
  if (x < a + 1.0) {
    // series: P = pre * sum_{k>=0} x^k / ((a+1)...(a+k))
    double term = 1.0, sum = 1.0, ak = a;
    for (int k = 0; k < 2000; ++k) {
      ak += 1.0;
      term *= x / ak;
      sum += term;
      if (term < sum * 1e-17) break;
    }
    P = pre * sum;
    P = P > 1.0 ? 1.0 : P;
    Q = 1.0 - P;
  } else {
    // continued fraction for Q (modified Lentz): Q = pre * a / (x + 1 - a - 1(1-a)/(x + 3 - a - ...))
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
      const double an = -(double)i * ((double)i - a);
      b += 2.0;
      d = an * d + b;
      if (fabs(d) < tiny) d = tiny;
      c = b + an / c;
      if (fabs(c) < tiny) c = tiny;
      d = 1.0 / d;
      const double del = d * c;
      h *= del;
      if (fabs(del - 1.0) < 1e-16) break;
    }
    Q = pre * a * h;
    Q = Q > 1.0 ? 1.0 : Q;
    P = 1.0 - Q;
  }
}

// x with P(a, x) = p (q = 1 - p given separately)
MCM_DS_COLD double gamma_inv(double a, double p, double q) {
  if (!(a > 0.0) || !(p >= 0.0) || !(q >= 0.0)) return __builtin_nan("");
  if (p == 0.0) return 0.0;
  if (q == 0.0) return __builtin_huge_val();
  // starting value: Wilson-Hilferty for a > 1, small-x / large-x expansions otherwise (Numerical Recipes invgammp)
  double x;
  const double gln = lgamma(a);
  if (a > 1.0) {
    const double z = norminv::norminvcdf(p < q ? p : q) * (p < q ? 1.0 : -1.0);   // Phi^-1(p)
    const double w = 1.0 - 1.0 / (9.0 * a) + z / (3.0 * sqrt(a));
    x = a * w * w * w;
    if (!(x > 0.0) || w <= 0.0) x = exp((log(p) + gln + log(a)) / a);     // far lower tail: leading series term
    if (x < 1e-3 * a) {
      const double x2 = exp((log(p) + gln + log(a)) / a);
      x = x2 < x || !(x > 0.0) ? x2 : x;
    }
  } else {
    const double t = 1.0 - a * (0.253 + a * 0.12);
    if (p < t) x = pow(p / t, 1.0 / a);
    else x = 1.0 - log1p(-(p - t) / (1.0 - t));
  }
  if (!(x > 0.0)) x = 1e-300;
  const bool lower = p <= q;
  const double am1 = a - 1.0;
  double lo = 0.0, hi = __builtin_huge_val();
  for (int it = 0; it < 60; ++it) {
    double P, Q;
    gamma_pq(a, x, P, Q);
    const double f = lower ? (P - p) : (q - Q);                    // both are "P(x) - p", formed without cancellation
    if (f > 0.0) hi = x; else lo = x;
    const double lpdf = am1 * log(x) - x - gln;
    const double pdf = exp(lpdf);
    if (!(pdf > 0.0)) { x = (hi < __builtin_huge_val()) ? 0.5 * (lo + hi) : 2.0 * x; continue; }
    const double u = f / pdf;
    double dx = u / (1.0 - 0.5 * fmin(1.0, fmax(-1.0, u * (am1 / x - 1.0))));      // Halley
    double xn = x - dx;
    if (!(xn > lo) || !(xn < hi)) xn = (hi < __builtin_huge_val()) ? 0.5 * (lo + hi) : 2.0 * x;   // keep the bracket
    const double step = fabs(xn - x);
    x = xn;
    if (step <= 2e-16 * x) break;
  }
  return x;
}

// ---- regularised incomplete beta I_x(a, b) -----------------------------------------------------------------------
MCM_DS_HD double beta_cf(double a, double b, double x) {
  const double tiny = 1e-300;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m < 3000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  return h;
}
// I_x(a,b) and its complement; y = 1 - x is passed separately so that x next to 1 keeps its accuracy
MCM_DS_HD void beta_pq(double a, double b, double x, double y, double& P, double& Q) {
  if (!(x > 0.0)) { P = 0.0; Q = 1.0; return; }
  if (!(y > 0.0)) { P = 1.0; Q = 0.0; return; }
  const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
  const double front = exp(a * log(x) + b * log(y) - lbeta);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    P = front * beta_cf(a, b, x) / a;
    P = P > 1.0 ? 1.0 : P;
    Q = 1.0 - P;
  } else {
    Q = front * beta_cf(b, a, y) / b;
    Q = Q > 1.0 ? 1.0 : Q;
    P = 1.0 - Q;
  }
}
// x with I_x(a,b) = p; also returns y = 1 - x with full relative accuracy
MCM_DS_COLD double beta_inv(double a, double b, double p, double q, double* y_out) {
  if (!(a > 0.0) || !(b > 0.0) || !(p >= 0.0) || !(q >= 0.0)) { if (y_out) *y_out = __builtin_nan(""); return __builtin_nan(""); }
  if (p == 0.0) { if (y_out) *y_out = 1.0; return 0.0; }
  if (q == 0.0) { if (y_out) *y_out = 0.0; return 1.0; }
  // solve for the smaller tail: if p > q swap roles (I_x(a,b) = 1 - I_{1-x}(b,a))
  const bool swap = p > q;
  const double aa = swap ? b : a, bb = swap ? a : b, pp = swap ? q : p;
  // starting value (Numerical Recipes invbetai)
  double x;
  if (aa >= 1.0 && bb >= 1.0) {
    const double t = sqrt(-2.0 * log(pp));
    double z = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
    const double al = (z * z - 3.0) / 6.0;
    const double hh = 2.0 / (1.0 / (2.0 * aa - 1.0) + 1.0 / (2.0 * bb - 1.0));
    const double w = (z * sqrt(al + hh) / hh) - (1.0 / (2.0 * bb - 1.0) - 1.0 / (2.0 * aa - 1.0)) * (al + 5.0 / 6.0 - 2.0 / (3.0 * hh));
    x = aa / (aa + bb * exp(2.0 * w));
  } else {
    const double lna = log(aa / (aa + bb)), lnb = log(bb / (aa + bb));
    const double t = exp(aa * lna) / aa, u = exp(bb * lnb) / bb;
    const double w = t + u;
    if (pp < t / w) x = pow(aa * w * pp, 1.0 / aa);
    else x = 1.0 - pow(bb * w * (1.0 - pp), 1.0 / bb);
  }
  if (!(x > 0.0)) x = 1e-300;
  if (!(x < 1.0)) x = 1.0 - 1e-16;
  const double lbeta = lgamma(aa) + lgamma(bb) - lgamma(aa + bb);
  double lo = 0.0, hi = 1.0;
  double y = 1.0 - x;
  for (int it = 0; it < 80; ++it) {
    double P, Q;
    beta_pq(aa, bb, x, y, P, Q);
    const double f = P - pp;                                         // pp <= 1/2: no cancellation trouble on this side
    if (f > 0.0) hi = x; else lo = x;
    const double lpdf = (aa - 1.0) * log(x) + (bb - 1.0) * log(y) - lbeta;
    const double pdf = exp(lpdf);
    if (!(pdf > 0.0) || !(pdf < 1e300)) { x = 0.5 * (lo + hi); y = 1.0 - x; continue; }
    const double u = f / pdf;
    double dx = u / (1.0 - 0.5 * fmin(1.0, fmax(-1.0, u * ((aa - 1.0) / x - (bb - 1.0) / y))));
    double xn = x - dx;
    if (!(xn > lo) || !(xn < hi)) { xn = 0.5 * (lo + hi); dx = x - xn; }
    const double step = fabs(dx);
    // y = 1 - x updated by the same step: keeps relative accuracy when x -> 1
    y = (xn > 0.5) ? (y + dx) : (1.0 - xn);
    if (!(y > 0.0)) y = 1.0 - xn;
    x = xn;
    if (step <= 2e-16 * fmin(x, y)) break;
  }
  if (swap) { if (y_out) *y_out = x; return y; }
  if (y_out) *y_out = y;
  return x;
}

// ---- quantile(d, p) ------------------------------------------------------------------------------------------------
// p in [0,1] and q = 1 - p (exact where the caller can provide it).  Parameter meaning per family = Distributions.jl.
MCM_DS_HD double quantile(int kind, const Params& d, double p, double q) {
  const double nan = __builtin_nan("");
  if (!(p >= 0.0 && p <= 1.0)) return nan;
  const double lq = (q > 0.5) ? log1p(-p) : log(q);      // log(1 - p)
  const double lp = (p > 0.5) ? log1p(-q) : log(p);      // log(p)
  switch (kind) {
    case NORMAL: return d.a + d.b * (p <= q ? norminv::norminvcdf(p) : -norminv::norminvcdf(q));   // mu + sigma Phi^-1(p)
    case UNIFORM: return d.a + (d.b - d.a) * p;
    case EXPONENTIAL: return -d.a * lq;                                                   // theta: scale
    case LOGNORMAL: return exp(d.a + d.b * (p <= q ? norminv::norminvcdf(p) : -norminv::norminvcdf(q)));
    case CAUCHY: {                                                                          // mu + sigma tan(pi (p - 1/2))
      if (p == 0.0) return -__builtin_huge_val();
      if (q == 0.0) return __builtin_huge_val();
      // tan(pi(p - 1/2)) = -cot(pi p) = cot(pi q): use the smaller of p, q for accuracy in the tails
      const double t = (p < q) ? -1.0 / tan(3.14159265358979323846 * p) : 1.0 / tan(3.14159265358979323846 * q);
      return d.a + d.b * t;
    }
    case LAPLACE: return p < q ? d.a + d.b * (lp + 0.6931471805599453) : d.a - d.b * (lq + 0.6931471805599453);
    case LOGISTIC: return d.a + d.b * (lp - lq);
    case RAYLEIGH: return d.a * sqrt(-2.0 * lq);
    case WEIBULL: return d.b * pow(-lq, 1.0 / d.a);                                       // alpha shape, theta scale
    case PARETO: return d.b * exp(-lq / d.a);                                             // alpha shape, theta scale
    case GUMBEL: return d.a - d.b * log(-lp);
    case TRIANGULAR: {                                                                     // a <= c <= b : d.a, d.b, d.c = a, b, c
      const double a = d.a, b = d.b, c = d.c;
      const double fc = (c - a) / (b - a);
      return p <= fc ? a + sqrt(p * (b - a) * (c - a)) : b - sqrt(q * (b - a) * (b - c));
    }
    case FRECHET: return d.b * pow(-lp, -1.0 / d.a);                                      // alpha shape, theta scale
    case KUMARASWAMY: return pow(-expm1(lq / d.b), 1.0 / d.a);                            // (1 - (1-p)^(1/b))^(1/a)
    case GAMMA: return d.b * gamma_inv(d.a, p, q);
    case CHISQ: return 2.0 * gamma_inv(0.5 * d.a, p, q);
    case INVERSEGAMMA: return d.b / gamma_inv(d.a, q, p);
    case BETA: return beta_inv(d.a, d.b, p, q, nullptr);
    case TDIST: {
      // |t| = sqrt(nu (1 - x)/x) with x = I^-1(2 min(p,q); nu/2, 1/2); both x and 1 - x come back accurately
      const double nu = d.a;
      const double pm = p < q ? p : q;
      if (pm == 0.0) return p < q ? -__builtin_huge_val() : __builtin_huge_val();
      double y;
      const double two_pm = 2.0 * pm;
      const double x = beta_inv(0.5 * nu, 0.5, two_pm, (p < q) ? (q - p) : (p - q), &y);
      const double t = sqrt(nu * y / x);
      return p < q ? -t : t;
    }
    case FDIST: {                                                                          // x = I^-1(p; nu1/2, nu2/2); F = nu2 x / (nu1 (1 - x))
      double y;
      const double x = beta_inv(0.5 * d.a, 0.5 * d.b, p, q, &y);
      return d.b * x / (d.a * y);
    }
    default: return nan;
  }
}

}  // namespace dist
}  // namespace mcm
