// mcm_quantile.cuh — host-side pieces of Statistics.quantile (type 7) shared by the single and
// the batched order-statistics paths.  [upstream Statistics.jl `_quantile`: h = (n-1)p + 1 evaluated
// as n*p + (1-p); j = clamp(trunc(h), 1, n-1); gamma = h - j; a = v[j], b = v[j+1];
// finite: a + gamma*(b-a), otherwise (1-gamma)*a + gamma*b]   (src/particles.jl:321 forwards to it)
#pragma once
#include <math.h>
#include "mcm_common.cuh"

namespace mcm {

// 0-based ranks of the two order statistics and the interpolation weight for probability p
inline void quantile_rank_pair(size_t n, double p, unsigned long long& r0, unsigned long long& r1, double& gamma) {
  const double aleph = (double)n * p + (1.0 - p);
  long long j = (long long)trunc(aleph);
  if (j > (long long)n - 1) j = (long long)n - 1;
  if (j < 1) j = 1;
  double g = aleph - (double)j;
  gamma = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
  if (n == 1) { r0 = 0; r1 = 0; }
  else { r0 = (unsigned long long)(j - 1); r1 = (unsigned long long)j; }
}

inline double quantile_combine(int dtype, double va, double vb, double g) {
  if (dtype == MCM_F64) {
    return (isfinite(va) && isfinite(vb)) ? va + g * (vb - va) : (1.0 - g) * va + g * vb;
  }
  // Float32 data, Float64 probability: (b-a) rounds in Float32, the rest promotes
  const float a = (float)va, b = (float)vb;
  return (isfinite(a) && isfinite(b)) ? (double)a + g * (double)(b - a) : (1.0 - g) * (double)a + g * (double)b;
}

}  // namespace mcm
