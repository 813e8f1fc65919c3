// Host twin of csrc/mcm_dist.cuh (TEST INFRASTRUCTURE): the same source compiled as plain C++ so that the inverse
// CDFs can be checked against scipy on the CPU.  Never linked into the product library.
#define MCM_HOST_TWIN 1
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_dist.cuh"
extern "C" {
void twin_dist_quantile(int kind, double a, double b, double c, const double* p, const double* q, double* out, long n) {
  mcm::dist::Params d{a, b, c};
  for (long i = 0; i < n; ++i) out[i] = mcm::dist::quantile(kind, d, p[i], q[i]);
}
}
