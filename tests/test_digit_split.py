"""The one-addition digit split of the Ozaki slicers (csrc/mcm_digits.cuh, used by k_slice / k_pslice / k_pz_coef) against
the digit-peeling loop it replaced: the SAME header compiled for the host, every digit of 6-, 7- and 8-digit splits over
random and extreme 64-bit integers.  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include <cstdio>
#include <cstdint>
#include <random>
#include "mcm_digits.cuh"
template <int S> static bool check(long long I0) {
  int8_t dg[8]; long long I = I0;
  for (int s = S - 1; s >= 1; --s) { const long long d = ((I + 128) & 255) - 128; dg[s] = (int8_t)d; I = (I - d) >> 8; }
  dg[0] = (int8_t)I;
  const unsigned long long K = mcm::balanced_digit_bytes<S>(I0);
  for (int s = 0; s < S; ++s) if ((int8_t)((K >> (8 * (S - 1 - s))) & 255) != dg[s]) return false;
  return true;
}
int main() {
  std::mt19937_64 g(1); long bad = 0;
  for (long t = 0; t < 3000000; ++t) {
    const long long r = (long long)g();
    bad += !check<6>(r >> 16) + !check<7>(r >> 8) + !check<8>(r) + !check<6>(r >> 40) + !check<7>(r >> 7) + !check<6>(r >> 15);
  }
  const long long ex[] = {0, 1, -1, 127, 128, -128, -129, 1ll << 47, -(1ll << 47), (1ll << 47) - 1, 1ll << 55, -(1ll << 55),
                          (1ll << 55) - 1, INT64_MAX, INT64_MIN, 32639, 32640, -32640, -32641};
  for (long long v : ex) bad += !check<6>(v) + !check<7>(v) + !check<8>(v);
  std::printf("%ld\n", bad);
  return bad != 0;
}
"""


def test_one_addition_digit_split_equals_the_peeling_loop(tmp_path):
    src = tmp_path / "digits.cpp"
    src.write_text(SRC)
    exe = tmp_path / "digits"
    inc = os.path.join(ROOT, "montecarlomeasurements.jl_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout + out.stderr
