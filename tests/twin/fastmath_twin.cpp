// Host twin of csrc/mcm_fastmath.cuh and csrc/mcm_norminv.cuh (TEST INFRASTRUCTURE): the same source
// compiled as plain C++ (-ffp-contract=off, explicit fma) is bit-identical to the device code, so ulp
// errors can be measured on the CPU over millions of points.  Never linked into the product library.
#define MCM_HOST_TWIN 1
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_fastmath.cuh"
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_norminv.cuh"
#include <cmath>
extern "C" {
void twin_sin(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::fast::sin_f64(x[i]); }
// the three forms of the sin / cos kernel (fixed parity, selected operands, per-lane branch) on the same reduced arguments
long twin_sincos_forms_differ(const double* r, const double* rl, long n) {
  long bad = 0;
  for (long i = 0; i < n; ++i)
    for (int odd = 0; odd < 2; ++odd) {
      const double a = odd ? mcm::fast::sincos_kernel_t<true>(r[i], rl[i]) : mcm::fast::sincos_kernel_t<false>(r[i], rl[i]);
      const double b = mcm::fast::sincos_kernel(r[i], rl[i], odd);
      const double c = mcm::fast::sincos_branch(r[i], rl[i], odd + 2);
      bad += (memcmp(&a, &b, 8) != 0) + (memcmp(&a, &c, 8) != 0);
    }
  return bad;
}
void twin_cos(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::fast::cos_f64(x[i]); }
void twin_exp(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::fast::exp_f64(x[i]); }
void twin_log(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::fast::log_f64(x[i]); }
void twin_pow(const double* x, const double* e, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::fast::pow_f64(x[i], e[i]); }
void twin_rem2pi(const double* x, int mode, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::fast::rem2pi_f64(x[i], mode); }
void twin_norminv(const double* p, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = mcm::norminv::norminvcdf(p[i]); }
// max |y - f(x)| in ulps of the long-double reference (64-bit mantissa: resolves ~0.001 ulp)
static double ulp_of(long double t) { double d = (double)fabsl(t); if (d == 0) return 4.9e-324; return nextafter(d, INFINITY) - d; }
double twin_maxulp(int fn, const double* x, const double* e, long n, double* worst_x) {
  double worst = 0;
  for (long i = 0; i < n; ++i) {
    double y; long double t;
    switch (fn) {
      case 0: y = mcm::fast::sin_f64(x[i]); t = sinl((long double)x[i]); break;
      case 1: y = mcm::fast::cos_f64(x[i]); t = cosl((long double)x[i]); break;
      case 2: y = mcm::fast::exp_f64(x[i]); t = expl((long double)x[i]); break;
      case 3: y = mcm::fast::pow_f64(x[i], e[i]); t = powl((long double)x[i], (long double)e[i]); break;
      default: y = mcm::fast::log_f64(x[i]); t = logl((long double)x[i]); break;
    }
    double u = (double)(fabsl((long double)y - t) / ulp_of(t));
    if (u > worst) { worst = u; *worst_x = x[i]; }
  }
  return worst;
}
}
