// mcm_digits.cuh — the balanced base-256 digit split of the Ozaki slicers (gram_tc.cu, pgemm_tc.cu), plain C++ so that the
// host test (tests/test_digit_split.py) compiles the very same function and compares it with the digit-peeling loop.
#pragma once
#ifdef __CUDACC__
#define MCM_DIGITS_HD __host__ __device__ __forceinline__
#else
#define MCM_DIGITS_HD inline
#endif

namespace mcm {

// The S signed digits d_b in [-128, 127] with I = sum_b d_b 256^b.  Since
//     I + sum_b 128 * 256^b = sum_b (d_b + 128) 256^b     with every (d_b + 128) in [0, 255],
// the UNSIGNED bytes of I + C (C = S bytes of 0x80) are d_b + 128, and flipping bit 7 of each byte turns them into the
// int8 digits: two 64-bit operations for all digits instead of ~10 per digit for the peeling loop
// (d = ((I + 128) & 255) - 128; I = (I - d) >> 8).  The top byte wraps exactly like the (int8_t) cast of the loop's
// remainder, so the result equals the loop's for EVERY 64-bit I.
template <int S>
MCM_DIGITS_HD unsigned long long balanced_digit_bytes(long long I) {
  static_assert(S >= 1 && S <= 8, "at most eight digits per 64-bit value");
  constexpr unsigned long long C = 0x8080808080808080ull >> (8 * (8 - S));
  return ((unsigned long long)I + C) ^ C;               // byte b (b < S) = digit of weight 256^b, as int8
}

}  // namespace mcm
