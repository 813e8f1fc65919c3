// mcm_keys.cuh — order-preserving integer keys for floating-point particles.
// Order is Julia's `isless` on floats: -0.0 < +0.0, every NaN last (Base sort / Statistics.quantile).
#pragma once
#include "mcm_common.cuh"

namespace mcm {

__device__ __forceinline__ uint64_t to_key(double x, bool& nan) {
  if (x != x) { nan = true; return ~0ull; }
  const uint64_t u = (uint64_t)__double_as_longlong(x);
  return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ uint32_t to_key(float x, bool& nan) {
  if (x != x) { nan = true; return ~0u; }
  const uint32_t u = __float_as_uint(x);
  return u ^ ((u >> 31) ? ~0u : 0x80000000u);
}
__device__ __forceinline__ double from_key(uint64_t k) {
  if (k == ~0ull) return __longlong_as_double(0x7ff8000000000000ll);
  const uint64_t u = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ float from_key(uint32_t k) {
  if (k == ~0u) return __uint_as_float(0x7fc00000u);
  const uint32_t u = k ^ ((k >> 31) ? 0x80000000u : ~0u);
  return __uint_as_float(u);
}


template <typename T> struct KeyOf;
template <> struct KeyOf<double> { using type = uint64_t; static constexpr int bits = 64; };
template <> struct KeyOf<float> { using type = uint32_t; static constexpr int bits = 32; };

}  // namespace mcm
