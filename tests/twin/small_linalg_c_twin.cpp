// Host twin of csrc/mcm_small_linalg_c.cuh (complex per-thread algorithms): the same source compiled as plain C++
// (-ffp-contract=off) so it can be pinned against LAPACK / scipy on the CPU.  Test infrastructure only.
//   batch layout: nb matrices, row-major, re and im in separate arrays.
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_small_linalg_c.cuh"
using namespace mcm::sla;

template <int N>
static int run(int op, const double* Ar, const double* Ai, double* outr, double* outi, long nb) {
  for (long b = 0; b < nb; ++b) {
    CMat<N> a;
    for (int i = 0; i < N * N; ++i) a.a[i] = cmk(Ar[b * N * N + i], Ai ? Ai[b * N * N + i] : 0.0);
    switch (op) {
      case 0: cexpm(a); for (int i = 0; i < N * N; ++i) { outr[b * N * N + i] = a.a[i].re; outi[b * N * N + i] = a.a[i].im; } break;
      case 1: clogm(a); for (int i = 0; i < N * N; ++i) { outr[b * N * N + i] = a.a[i].re; outi[b * N * N + i] = a.a[i].im; } break;
      case 2: { int piv[N]; const int s = clu_factor(a, piv); const cx d = s == 0 ? cmk(0.0, 0.0) : cdet_from_lu(a, s); outr[b] = d.re; outi[b] = d.im; break; }
      case 3: { cx w[N]; ceigvals(a, w); for (int i = 0; i < N; ++i) { outr[b * N + i] = w[i].re; outi[b * N + i] = w[i].im; } break; }
      case 4: { double sv[N]; csvdvals(a, sv); for (int i = 0; i < N; ++i) outr[b * N + i] = sv[i]; break; }
      default: return 1;
    }
  }
  return 0;
}
extern "C" int slac_twin(int op, int n, const double* Ar, const double* Ai, double* outr, double* outi, long nb) {
  switch (n) {
    case 1: return run<1>(op, Ar, Ai, outr, outi, nb);
    case 2: return run<2>(op, Ar, Ai, outr, outi, nb);
    case 3: return run<3>(op, Ar, Ai, outr, outi, nb);
    case 4: return run<4>(op, Ar, Ai, outr, outi, nb);
    case 5: return run<5>(op, Ar, Ai, outr, outi, nb);
    case 6: return run<6>(op, Ar, Ai, outr, outi, nb);
    case 7: return run<7>(op, Ar, Ai, outr, outi, nb);
    case 8: return run<8>(op, Ar, Ai, outr, outi, nb);
    case 9: return run<9>(op, Ar, Ai, outr, outi, nb);
    case 10: return run<10>(op, Ar, Ai, outr, outi, nb);
    case 11: return run<11>(op, Ar, Ai, outr, outi, nb);
    case 12: return run<12>(op, Ar, Ai, outr, outi, nb);
    case 13: return run<13>(op, Ar, Ai, outr, outi, nb);
    case 14: return run<14>(op, Ar, Ai, outr, outi, nb);
    case 15: return run<15>(op, Ar, Ai, outr, outi, nb);
    case 16: return run<16>(op, Ar, Ai, outr, outi, nb);
    default: return 3;
  }
}
