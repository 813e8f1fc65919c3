// Host twin of csrc/mcm_small_linalg.cuh: the same source compiled as plain C++ (-ffp-contract=off) so the
// per-thread algorithms can be pinned against LAPACK / scipy on the CPU.  Test infrastructure only.
//   batch layout: nb matrices, row-major, contiguous.
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_small_linalg.cuh"
using namespace mcm::sla;

template <int N>
static int run(int op, const double* A, const double* B, int m, double* out, double* out_im, long nb) {
  for (long b = 0; b < nb; ++b) {
    Mat<N> a;
    for (int i = 0; i < N * N; ++i) a.a[i] = A[b * N * N + i];
    switch (op) {
      case 0: {  // solve: B is N x m
        int piv[N];
        lu_factor(a, piv);
        for (int j = 0; j < m; ++j) {
          double col[N];
          for (int i = 0; i < N; ++i) col[i] = B[(b * N + i) * m + j];
          lu_solve(a, piv, col, 1);
          for (int i = 0; i < N; ++i) out[(b * N + i) * m + j] = col[i];
        }
        break;
      }
      case 1: {  // inverse
        int piv[N];
        lu_factor(a, piv);
        Mat<N> x;
        set_identity(x, 1.0);
        for (int j = 0; j < N; ++j) lu_solve(a, piv, &x.a[j], N);
        for (int i = 0; i < N * N; ++i) out[b * N * N + i] = x.a[i];
        break;
      }
      case 2: {
        int piv[N];
        const int s = lu_factor(a, piv);
        out[b] = s == 0 ? 0.0 : det_from_lu(a, s);
        break;
      }
      case 3: {
        expm(a);
        for (int i = 0; i < N * N; ++i) out[b * N * N + i] = a.a[i];
        break;
      }
      case 4: {
        double wr[N], wi[N];
        eig_general(a, wr, wi);
        for (int i = 0; i < N; ++i) { out[b * N + i] = wr[i]; out_im[b * N + i] = wi[i]; }
        break;
      }
      case 5: {
        double w[N];
        eig_sym(a, w);
        for (int i = 0; i < N; ++i) out[b * N + i] = w[i];
        break;
      }
      case 6: {
        if constexpr (N <= 8) {
          Mat<N> c, x;
          for (int i = 0; i < N * N; ++i) c.a[i] = B[b * N * N + i];
          lyap(a, c, x);
          for (int i = 0; i < N * N; ++i) out[b * N * N + i] = x.a[i];
        } else {
          return 2;
        }
        break;
      }
      case 7: {
        double sv[N];
        svdvals(a, sv);
        for (int i = 0; i < N; ++i) out[b * N + i] = sv[i];
        break;
      }
      default: return 1;
    }
  }
  return 0;
}

extern "C" int sla_twin(int op, int n, const double* A, const double* B, int m, double* out, double* out_im, long nb) {
  switch (n) {
    case 1: return run<1>(op, A, B, m, out, out_im, nb);
    case 2: return run<2>(op, A, B, m, out, out_im, nb);
    case 3: return run<3>(op, A, B, m, out, out_im, nb);
    case 4: return run<4>(op, A, B, m, out, out_im, nb);
    case 5: return run<5>(op, A, B, m, out, out_im, nb);
    case 6: return run<6>(op, A, B, m, out, out_im, nb);
    case 7: return run<7>(op, A, B, m, out, out_im, nb);
    case 8: return run<8>(op, A, B, m, out, out_im, nb);
    case 9: return run<9>(op, A, B, m, out, out_im, nb);
    case 10: return run<10>(op, A, B, m, out, out_im, nb);
    case 11: return run<11>(op, A, B, m, out, out_im, nb);
    case 12: return run<12>(op, A, B, m, out, out_im, nb);
    case 13: return run<13>(op, A, B, m, out, out_im, nb);
    case 14: return run<14>(op, A, B, m, out, out_im, nb);
    case 15: return run<15>(op, A, B, m, out, out_im, nb);
    case 16: return run<16>(op, A, B, m, out, out_im, nb);
  }
  return 3;
}
