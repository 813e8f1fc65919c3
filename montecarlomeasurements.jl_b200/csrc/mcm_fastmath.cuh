// mcm_fastmath.cuh — Float64 kernels for the hot transcendental primitives of
// the headline chain (exp, sin, cos, log, pow).  These are the device
// counterparts of SLEEFPirates.exp / sin_fast / cos_fast / log_fast
// (src/sleefpirates.jl:2-36) and Base `^` (src/particles.jl:183).
//
// Round-1 baseline: CUDA's libdevice double-precision routines (documented
// maximum errors: exp 1 ulp, log 1 ulp, sin/cos 2 ulp, pow 2 ulp).
#pragma once
#include <math.h>

namespace mcm {
namespace fast {

__device__ __forceinline__ double exp_f64(double x) { return ::exp(x); }
__device__ __forceinline__ double sin_f64(double x) { return ::sin(x); }
__device__ __forceinline__ double cos_f64(double x) { return ::cos(x); }
__device__ __forceinline__ double log_f64(double x) { return ::log(x); }
__device__ __forceinline__ double pow_f64(double x, double y) { return ::pow(x, y); }

}  // namespace fast
}  // namespace mcm
