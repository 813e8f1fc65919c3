// mcm_small_linalg.cuh — dense kernels for ONE small matrix held by ONE thread.
// The reference runs `f(vecindex.(p, i))` on the host for every particle i (ℝⁿ2ℝⁿ_function / ℝⁿ2ℂⁿ_function,
// src/particles.jl:241-282, :395-406, :603-607); here particle i is thread i and the matrix lives in its registers /
// local memory.  Plain C++ (no CUDA types): the same source is compiled for the host by tests/twin to pin it against
// LAPACK on the CPU; the library is built with -fmad=false, so both give the same bits.
//   lu_factor / lu_solve   partial pivoting, first maximal pivot (LAPACK getrf order of operations, unblocked)
//   expm                   Higham (2005) scaling and squaring, Padé 3/5/7/9/13 by the 1-norm — the method of
//                          LinearAlgebra.exp! (no balancing step here)
//   eig_sym                cyclic Jacobi, ascending
//   eig_general            elimination to Hessenberg form + Francis double-shift QR (EISPACK hqr); returned sorted by
//                          (real, imaginary) — LAPACK's deflation order is not reproducible
//   lyap                   A X + X A' + C = 0 through the Kronecker system (n <= 8: order n^2 <= 64)
//   svdvals                one-sided Jacobi (Hestenes), descending
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MCM_SLA __device__ __forceinline__
// loops over N are fully unrolled up to N = 8 (everything in registers); above, the matrix lives in local memory and
// the loops stay rolled (same operations in the same order: the bits do not depend on it)
#define MCM_SLA_UNROLL _Pragma("unroll (N <= 8 ? 64 : 1)")
#else
#define MCM_SLA inline
#define MCM_SLA_UNROLL
#endif

namespace mcm {
namespace sla {

constexpr double kEps = 2.220446049250313e-16;

template <int N>
struct Mat {
  double a[N * N];
  MCM_SLA double& operator()(int i, int j) { return a[i * N + j]; }
  MCM_SLA const double& operator()(int i, int j) const { return a[i * N + j]; }
};

template <int N>
MCM_SLA void set_identity(Mat<N>& m, double v) {
  for (int i = 0; i < N * N; ++i) m.a[i] = 0.0;
  for (int i = 0; i < N; ++i) m(i, i) = v;
}
template <int N>
MCM_SLA void matmul(Mat<N>& c, const Mat<N>& a, const Mat<N>& b) {   // c must not alias a or b
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int k = 0; k < N; ++k) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
}
template <int N>
MCM_SLA double norm1(const Mat<N>& a) {
  double best = 0.0;
  for (int j = 0; j < N; ++j) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += fabs(a(i, j));
    best = s > best || s != s ? s : best;
  }
  return best;
}

// returns the permutation sign, 0 if a pivot is exactly zero (the factorisation is still completed as getrf does)
template <int N>
MCM_SLA int lu_factor(Mat<N>& a, int (&piv)[N]) {
  int sign = 1;
  bool singular = false;
  for (int k = 0; k < N; ++k) {
    int p = k;
    double mx = fabs(a(k, k));
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(a(i, k));
      if (v > mx) { mx = v; p = i; }
    }
    piv[k] = p;
    if (mx == 0.0) { singular = true; continue; }
    if (p != k) {
      for (int j = 0; j < N; ++j) { const double t = a(k, j); a(k, j) = a(p, j); a(p, j) = t; }
      sign = -sign;
    }
    const double inv = 1.0 / a(k, k);
    for (int i = k + 1; i < N; ++i) {
      const double l = a(i, k) * inv;
      a(i, k) = l;
      for (int j = k + 1; j < N; ++j) a(i, j) -= l * a(k, j);
    }
  }
  return singular ? 0 : sign;
}
// b: N values with stride `ldb`
template <int N>
MCM_SLA void lu_solve(const Mat<N>& lu, const int (&piv)[N], double* b, int ldb) {
  for (int k = 0; k < N; ++k) {
    const int p = piv[k];
    if (p != k) { const double t = b[k * ldb]; b[k * ldb] = b[p * ldb]; b[p * ldb] = t; }
  }
  for (int i = 1; i < N; ++i) {
    double s = b[i * ldb];
    for (int k = 0; k < i; ++k) s -= lu(i, k) * b[k * ldb];
    b[i * ldb] = s;
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i * ldb];
    for (int k = i + 1; k < N; ++k) s -= lu(i, k) * b[k * ldb];
    b[i * ldb] = s / lu(i, i);
  }
}
template <int N>
MCM_SLA double det_from_lu(const Mat<N>& lu, int sign) {
  double d = (double)sign;
  for (int i = 0; i < N; ++i) d *= lu(i, i);
  return d;
}

// x <- exp(x); returns false if the Padé denominator was singular (non-finite input)
template <int N>
MCM_SLA bool expm(Mat<N>& x) {
  const double nA = norm1(x);
  Mat<N> a2, u, v, p, t;
  matmul(a2, x, x);
  int squarings = 0;
  if (nA <= 2.1) {
    const double c9[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    const double c7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    const double c5[6] = {30240., 15120., 3360., 420., 30., 1.};
    const double c3[4] = {120., 60., 12., 1.};
    const double* c = nA > 0.95 ? c9 : (nA > 0.25 ? c7 : (nA > 0.015 ? c5 : c3));
    const int nc = nA > 0.95 ? 10 : (nA > 0.25 ? 8 : (nA > 0.015 ? 6 : 4));
    set_identity(p, 1.0);
    set_identity(u, c[1]);
    set_identity(v, c[0]);
    for (int k = 1; k <= nc / 2 - 1; ++k) {
      matmul(t, p, a2);
      p = t;
      for (int i = 0; i < N * N; ++i) { u.a[i] += c[2 * k + 1] * p.a[i]; v.a[i] += c[2 * k] * p.a[i]; }
    }
    matmul(t, x, u);
    u = t;
  } else {
    const double s = log2(nA / 5.4);
    if (s > 0.0) {
      squarings = (int)ceil(s);
      const double f = ldexp(1.0, -squarings);
      for (int i = 0; i < N * N; ++i) x.a[i] *= f;
      matmul(a2, x, x);
    }
    const double c[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                          129060195264000., 10559470521600., 670442572800., 33522128640.,
                          1323241920., 40840800., 960960., 16380., 182., 1.};
    Mat<N> a4, a6;
    matmul(a4, a2, a2);
    matmul(a6, a2, a4);
    for (int i = 0; i < N * N; ++i) t.a[i] = c[13] * a6.a[i] + c[11] * a4.a[i] + c[9] * a2.a[i];
    matmul(p, a6, t);
    for (int i = 0; i < N * N; ++i) p.a[i] += c[7] * a6.a[i] + c[5] * a4.a[i] + c[3] * a2.a[i];
    for (int i = 0; i < N; ++i) p(i, i) += c[1];
    matmul(u, x, p);
    for (int i = 0; i < N * N; ++i) t.a[i] = c[12] * a6.a[i] + c[10] * a4.a[i] + c[8] * a2.a[i];
    matmul(v, a6, t);
    for (int i = 0; i < N * N; ++i) v.a[i] += c[6] * a6.a[i] + c[4] * a4.a[i] + c[2] * a2.a[i];
    for (int i = 0; i < N; ++i) v(i, i) += c[0];
  }
  // X = (V - U) \ (V + U)
  for (int i = 0; i < N * N; ++i) { p.a[i] = v.a[i] - u.a[i]; x.a[i] = v.a[i] + u.a[i]; }
  int piv[N];
  const int sign = lu_factor(p, piv);
  for (int j = 0; j < N; ++j) lu_solve(p, piv, &x.a[j], N);
  for (int k = 0; k < squarings; ++k) {
    matmul(t, x, x);
    x = t;
  }
  return sign != 0;
}

template <int N>
MCM_SLA void sort_ascending(double (&w)[N]) {
  for (int i = 1; i < N; ++i) {
    const double v = w[i];
    int j = i - 1;
    while (j >= 0 && w[j] > v) { w[j + 1] = w[j]; --j; }
    w[j + 1] = v;
  }
}

// symmetric eigenvalues (the lower triangle is taken as given by the upper one), ascending
template <int N>
MCM_SLA void eig_sym(Mat<N>& a, double (&w)[N]) {
  MCM_SLA_UNROLL
  for (int i = 0; i < N; ++i)
    MCM_SLA_UNROLL
    for (int j = i + 1; j < N; ++j) a(j, i) = a(i, j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    MCM_SLA_UNROLL
    for (int i = 0; i < N; ++i) {
      diag += a(i, i) * a(i, i);
      MCM_SLA_UNROLL
      for (int j = i + 1; j < N; ++j) off += a(i, j) * a(i, j);
    }
    if (!(off > 1e-34 * diag) || off == 0.0) break;
    // fully unrolled: every index is a compile-time constant, so the matrix stays in registers on the device
    MCM_SLA_UNROLL
    for (int p = 0; p < N - 1; ++p)
      MCM_SLA_UNROLL
      for (int q = p + 1; q < N; ++q) {
        const double apq = a(p, q);
        if (apq == 0.0) continue;
        const double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        MCM_SLA_UNROLL
        for (int k = 0; k < N; ++k) {          // columns p, q
          const double akp = a(k, p), akq = a(k, q);
          a(k, p) = c * akp - s * akq;
          a(k, q) = s * akp + c * akq;
        }
        MCM_SLA_UNROLL
        for (int k = 0; k < N; ++k) {          // rows p, q
          const double apk = a(p, k), aqk = a(q, k);
          a(p, k) = c * apk - s * aqk;
          a(q, k) = s * apk + c * aqk;
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = a(i, i);
  sort_ascending(w);
}

MCM_SLA double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// general real eigenvalues; returns false if the QR iteration did not converge (values are then NaN)
template <int N>
MCM_SLA bool eig_general(Mat<N>& a, double (&wr)[N], double (&wi)[N]) {
  // -- reduction to upper Hessenberg form by stabilised elementary similarity transformations
  for (int m = 1; m < N - 1; ++m) {
    double x = 0.0;
    int i = m;
    for (int j = m; j < N; ++j)
      if (fabs(a(j, m - 1)) > fabs(x)) { x = a(j, m - 1); i = j; }
    if (i != m) {
      for (int j = m - 1; j < N; ++j) { const double t = a(i, j); a(i, j) = a(m, j); a(m, j) = t; }
      for (int j = 0; j < N; ++j) { const double t = a(j, i); a(j, i) = a(j, m); a(j, m) = t; }
    }
    if (x != 0.0) {
      for (i = m + 1; i < N; ++i) {
        double y = a(i, m - 1);
        if (y != 0.0) {
          y /= x;
          a(i, m - 1) = y;
          for (int j = m; j < N; ++j) a(i, j) -= y * a(m, j);
          for (int j = 0; j < N; ++j) a(j, m) += y * a(j, i);
        }
      }
    }
  }
  for (int i = 2; i < N; ++i)
    for (int j = 0; j < i - 1; ++j) a(i, j) = 0.0;
  // -- Francis double-shift QR on the Hessenberg matrix
  double anorm = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = (i > 0 ? i - 1 : 0); j < N; ++j) anorm += fabs(a(i, j));
  int nn = N - 1;
  double t = 0.0;
  bool ok = true;
  while (nn >= 0 && ok) {
    int its = 0, l;
    do {
      for (l = nn; l > 0; --l) {
        double s = fabs(a(l - 1, l - 1)) + fabs(a(l, l));
        if (s == 0.0) s = anorm;
        if (fabs(a(l, l - 1)) <= kEps * s) { a(l, l - 1) = 0.0; break; }
      }
      double x = a(nn, nn);
      if (l == nn) {
        wr[nn] = x + t; wi[nn] = 0.0; --nn;
      } else {
        double y = a(nn - 1, nn - 1);
        double w = a(nn, nn - 1) * a(nn - 1, nn);
        if (l == nn - 1) {
          const double p = 0.5 * (y - x);
          const double q = p * p + w;
          double z = sqrt(fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + sign_of(z, p);
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn - 1] = z; wi[nn] = -z;
          }
          nn -= 2;
        } else {
          if (its == 60) { ok = false; break; }
          if (its == 10 || its == 20 || its == 40) {      // exceptional shift
            t += x;
            for (int i = 0; i <= nn; ++i) a(i, i) -= x;
            const double s = fabs(a(nn, nn - 1)) + fabs(a(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          double p = 0.0, q = 0.0, r = 0.0, z;
          for (m = nn - 2; m >= l; --m) {
            z = a(m, m);
            r = x - z;
            double s = y - z;
            p = (r * s - w) / a(m + 1, m) + a(m, m + 1);
            q = a(m + 1, m + 1) - z - r - s;
            r = a(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            p /= s; q /= s; r /= s;
            if (m == l) break;
            const double u = fabs(a(m, m - 1)) * (fabs(q) + fabs(r));
            const double v = fabs(p) * (fabs(a(m - 1, m - 1)) + fabs(z) + fabs(a(m + 1, m + 1)));
            if (u <= kEps * v) break;
          }
          for (int i = m; i < nn - 1; ++i) {
            a(i + 2, i) = 0.0;
            if (i != m) a(i + 2, i - 1) = 0.0;
          }
          for (int k = m; k < nn; ++k) {
            if (k != m) {
              p = a(k, k - 1);
              q = a(k + 1, k - 1);
              r = 0.0;
              if (k + 1 != nn) r = a(k + 2, k - 1);
              x = fabs(p) + fabs(q) + fabs(r);
              if (x != 0.0) { p /= x; q /= x; r /= x; }
            }
            const double s = sign_of(sqrt(p * p + q * q + r * r), p);
            if (s != 0.0) {
              if (k == m) {
                if (l != m) a(k, k - 1) = -a(k, k - 1);
              } else {
                a(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s; y = q / s; z = r / s;
              q /= p; r /= p;
              for (int j = k; j <= nn; ++j) {
                p = a(k, j) + q * a(k + 1, j);
                if (k + 1 != nn) { p += r * a(k + 2, j); a(k + 2, j) -= p * z; }
                a(k + 1, j) -= p * y;
                a(k, j) -= p * x;
              }
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {
                p = x * a(i, k) + y * a(i, k + 1);
                if (k + 1 != nn) { p += z * a(i, k + 2); a(i, k + 2) -= p * r; }
                a(i, k + 1) -= p * q;
                a(i, k) -= p;
              }
            }
          }
        }
      }
    } while (ok && l + 1 < nn);
  }
  if (!ok) {
    for (int i = 0; i < N; ++i) { wr[i] = wi[i] = nan(""); }
    return false;
  }
  // order: real part, then imaginary part
  for (int i = 1; i < N; ++i) {
    const double vr = wr[i], vi = wi[i];
    int j = i - 1;
    while (j >= 0 && (wr[j] > vr || (wr[j] == vr && wi[j] > vi))) { wr[j + 1] = wr[j]; wi[j + 1] = wi[j]; --j; }
    wr[j + 1] = vr; wi[j + 1] = vi;
  }
  return true;
}

// X with A X + X A' + C = 0; returns false if the Kronecker system is singular
template <int N>
MCM_SLA bool lyap(const Mat<N>& a, const Mat<N>& c, Mat<N>& x) {
  constexpr int K = N * N;
  Mat<K> k;
  for (int i = 0; i < K * K; ++i) k.a[i] = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      const int r = i * N + j;
      for (int q = 0; q < N; ++q) {
        k(r, q * N + j) += a(i, q);     // (A X)_ij
        k(r, i * N + q) += a(j, q);     // (X A')_ij
      }
      x.a[r] = -c.a[r];
    }
  int piv[K];
  const int sign = lu_factor(k, piv);
  lu_solve(k, piv, x.a, 1);
  return sign != 0;
}

// singular values, descending
template <int N>
MCM_SLA void svdvals(Mat<N>& a, double (&sv)[N]) {
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    MCM_SLA_UNROLL
    for (int p = 0; p < N - 1; ++p)
      MCM_SLA_UNROLL
      for (int q = p + 1; q < N; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        MCM_SLA_UNROLL
        for (int i = 0; i < N; ++i) {
          alpha += a(i, p) * a(i, p);
          beta += a(i, q) * a(i, q);
          gamma += a(i, p) * a(i, q);
        }
        if (gamma == 0.0 || !(fabs(gamma) > kEps * sqrt(alpha * beta))) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
        MCM_SLA_UNROLL
        for (int i = 0; i < N; ++i) {
          const double aip = a(i, p), aiq = a(i, q);
          a(i, p) = c * aip - s * aiq;
          a(i, q) = s * aip + c * aiq;
        }
      }
    if (!rotated) break;
  }
  for (int j = 0; j < N; ++j) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += a(i, j) * a(i, j);
    sv[j] = -sqrt(s);
  }
  sort_ascending(sv);
  for (int j = 0; j < N; ++j) sv[j] = -sv[j];
}

}  // namespace sla
}  // namespace mcm
