// mcm_small_linalg_c.cuh — dense kernels for ONE small COMPLEX matrix held by ONE thread.
// The reference evaluates exp / log / det / eigvals / svdvals of a Matrix{Complex{Particles}} — and log of a real
// uncertain matrix, whose result is complex — by looping over the particle index on the host and calling LAPACK on
// every sample (ℂⁿ2ℂⁿ_function, ℂⁿ2ℂ_function, complex.jl:104-181; Base.log(::Matrix{<:AbstractParticles}) =
// ℝⁿ2ℂⁿ_function(log, p), particles.jl:603).  Here particle i is thread i.  Plain C++ (no CUDA types, no <complex>): the
// same source is compiled for the host by tests/twin and pinned against LAPACK / scipy; -fmad=false on both sides.
//   clu_factor / cdet     partial pivoting on |re| + |im| (LAPACK izamax), unblocked getf2 order
//   cexpm                 Higham (2005) scaling and squaring, Padé 3/5/7/9/13 by the 1-norm (LinearAlgebra.exp!)
//   cschur                Householder reduction to Hessenberg form + single-shift QR with Givens rotations (Wilkinson
//                         shift, exceptional shifts), A = Q T Q^H with T upper triangular; ceigvals = diag(T) ordered by
//                         (re, im) — LAPACK's deflation order is not reproducible
//   clogm                 the method of LinearAlgebra.log: Schur form, inverse scaling and squaring on the triangular
//                         factor (square roots by the Björck-Hammarling recurrence until ||T - I||_1 <= 0.264, then the
//                         [7/7] Padé approximant of log(I + X) in partial fractions at the Gauss-Legendre nodes, Higham
//                         2008 alg. 11.9), back-transformation.  Principal branch; eigenvalues on the negative real axis
//                         give Im = +pi like Base
//   csvdvals              one-sided Jacobi (Hestenes) with the column pair's phase removed first, descending
#pragma once
#include <math.h>
#include "mcm_small_linalg.cuh"

namespace mcm {
namespace sla {

struct cx { double re, im; };
MCM_SLA cx cmk(double re, double im) { cx z; z.re = re; z.im = im; return z; }
MCM_SLA cx cadd(cx a, cx b) { return cmk(a.re + b.re, a.im + b.im); }
MCM_SLA cx csub(cx a, cx b) { return cmk(a.re - b.re, a.im - b.im); }
MCM_SLA cx cmul(cx a, cx b) { return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
MCM_SLA cx cscale(cx a, double s) { return cmk(a.re * s, a.im * s); }
MCM_SLA cx cconj(cx a) { return cmk(a.re, -a.im); }
MCM_SLA double cabs1(cx a) { return fabs(a.re) + fabs(a.im); }
MCM_SLA double cabs2(cx a) { return a.re * a.re + a.im * a.im; }
// |a| from IEEE operations only (libm's hypot differs between glibc and libdevice in the last bit, and the host twin
// must reproduce the device bit for bit); scaled by the larger component: no overflow, <= 2 ulp
MCM_SLA double hyp(double x, double y) {
  x = fabs(x); y = fabs(y);
  const double m = x > y ? x : y, l = x > y ? y : x;
  if (m == 0.0 || m != m || l != l) return m + l;
  if (m > 1.7976931348623157e308) return m;
  const double r = l / m;
  return m * sqrt(1.0 + r * r);
}
MCM_SLA double cabsv(cx a) { return hyp(a.re, a.im); }
// Smith's algorithm: no overflow in the intermediate products
MCM_SLA cx cdiv(cx a, cx b) {
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, d = b.re + b.im * r;
    return cmk((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  }
  const double r = b.re / b.im, d = b.re * r + b.im;
  return cmk((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
// principal square root (Re >= 0; a negative real number gives +i sqrt|x|)
MCM_SLA cx csqrtp(cx z) {
  const double m = cabsv(z);
  if (m == 0.0) return cmk(0.0, 0.0);
  if (z.re >= 0.0) {
    const double t = sqrt(0.5 * (m + z.re));
    return cmk(t, z.im / (2.0 * t));
  }
  const double t = sqrt(0.5 * (m - z.re));
  return cmk(fabs(z.im) / (2.0 * t), z.im < 0.0 ? -t : t);
}

template <int N>
struct CMat {
  cx a[N * N];
  MCM_SLA cx& operator()(int i, int j) { return a[i * N + j]; }
  MCM_SLA const cx& operator()(int i, int j) const { return a[i * N + j]; }
};
template <int N>
MCM_SLA void cset_identity(CMat<N>& m, double v) {
  for (int i = 0; i < N * N; ++i) m.a[i] = cmk(0.0, 0.0);
  for (int i = 0; i < N; ++i) m(i, i) = cmk(v, 0.0);
}
template <int N>
MCM_SLA void cmatmul(CMat<N>& c, const CMat<N>& a, const CMat<N>& b) {   // c must not alias a or b
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      cx s = cmk(0.0, 0.0);
      for (int k = 0; k < N; ++k) s = cadd(s, cmul(a(i, k), b(k, j)));
      c(i, j) = s;
    }
}
template <int N>
MCM_SLA double cnorm1(const CMat<N>& a) {
  double best = 0.0;
  for (int j = 0; j < N; ++j) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += cabsv(a(i, j));
    best = s > best || s != s ? s : best;
  }
  return best;
}

template <int N>
MCM_SLA int clu_factor(CMat<N>& a, int (&piv)[N]) {
  int sign = 1;
  bool singular = false;
  for (int k = 0; k < N; ++k) {
    int p = k;
    double mx = cabs1(a(k, k));
    for (int i = k + 1; i < N; ++i) {
      const double v = cabs1(a(i, k));
      if (v > mx) { mx = v; p = i; }
    }
    piv[k] = p;
    if (mx == 0.0 || mx != mx) { singular = true; continue; }
    if (p != k) {
      for (int j = 0; j < N; ++j) { const cx t = a(k, j); a(k, j) = a(p, j); a(p, j) = t; }
      sign = -sign;
    }
    const cx inv = cdiv(cmk(1.0, 0.0), a(k, k));
    for (int i = k + 1; i < N; ++i) {
      const cx l = cmul(a(i, k), inv);
      a(i, k) = l;
      for (int j = k + 1; j < N; ++j) a(i, j) = csub(a(i, j), cmul(l, a(k, j)));
    }
  }
  return singular ? 0 : sign;
}
template <int N>
MCM_SLA void clu_solve(const CMat<N>& lu, const int (&piv)[N], cx* b, int ldb) {
  for (int k = 0; k < N; ++k) {
    const int p = piv[k];
    if (p != k) { const cx t = b[k * ldb]; b[k * ldb] = b[p * ldb]; b[p * ldb] = t; }
  }
  for (int i = 1; i < N; ++i) {
    cx s = b[i * ldb];
    for (int k = 0; k < i; ++k) s = csub(s, cmul(lu(i, k), b[k * ldb]));
    b[i * ldb] = s;
  }
  for (int i = N - 1; i >= 0; --i) {
    cx s = b[i * ldb];
    for (int k = i + 1; k < N; ++k) s = csub(s, cmul(lu(i, k), b[k * ldb]));
    b[i * ldb] = cdiv(s, lu(i, i));
  }
}
template <int N>
MCM_SLA cx cdet_from_lu(const CMat<N>& lu, int sign) {
  cx d = cmk((double)sign, 0.0);
  for (int i = 0; i < N; ++i) d = cmul(d, lu(i, i));
  return d;
}

// x <- exp(x); returns false if the Padé denominator was singular (non-finite input)
template <int N>
MCM_SLA bool cexpm(CMat<N>& x) {
  const double nA = cnorm1(x);
  CMat<N> a2, u, v, p, t;
  cmatmul(a2, x, x);
  int squarings = 0;
  if (nA <= 2.1) {
    const double c9[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    const double c7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    const double c5[6] = {30240., 15120., 3360., 420., 30., 1.};
    const double c3[4] = {120., 60., 12., 1.};
    const double* c = nA > 0.95 ? c9 : (nA > 0.25 ? c7 : (nA > 0.015 ? c5 : c3));
    const int nc = nA > 0.95 ? 10 : (nA > 0.25 ? 8 : (nA > 0.015 ? 6 : 4));
    cset_identity(p, 1.0);
    cset_identity(u, c[1]);
    cset_identity(v, c[0]);
    for (int k = 1; k <= nc / 2 - 1; ++k) {
      cmatmul(t, p, a2);
      p = t;
      for (int i = 0; i < N * N; ++i) {
        u.a[i] = cadd(u.a[i], cscale(p.a[i], c[2 * k + 1]));
        v.a[i] = cadd(v.a[i], cscale(p.a[i], c[2 * k]));
      }
    }
    cmatmul(t, x, u);
    u = t;
  } else {
    const double s = log2(nA / 5.4);
    if (s > 0.0) {
      squarings = (int)ceil(s);
      const double f = ldexp(1.0, -squarings);
      for (int i = 0; i < N * N; ++i) x.a[i] = cscale(x.a[i], f);
      cmatmul(a2, x, x);
    }
    const double c[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                          129060195264000., 10559470521600., 670442572800., 33522128640.,
                          1323241920., 40840800., 960960., 16380., 182., 1.};
    CMat<N> a4, a6;
    cmatmul(a4, a2, a2);
    cmatmul(a6, a2, a4);
    for (int i = 0; i < N * N; ++i) t.a[i] = cadd(cadd(cscale(a6.a[i], c[13]), cscale(a4.a[i], c[11])), cscale(a2.a[i], c[9]));
    cmatmul(p, a6, t);
    for (int i = 0; i < N * N; ++i) p.a[i] = cadd(p.a[i], cadd(cadd(cscale(a6.a[i], c[7]), cscale(a4.a[i], c[5])), cscale(a2.a[i], c[3])));
    for (int i = 0; i < N; ++i) p(i, i).re += c[1];
    cmatmul(u, x, p);
    for (int i = 0; i < N * N; ++i) t.a[i] = cadd(cadd(cscale(a6.a[i], c[12]), cscale(a4.a[i], c[10])), cscale(a2.a[i], c[8]));
    cmatmul(v, a6, t);
    for (int i = 0; i < N * N; ++i) v.a[i] = cadd(v.a[i], cadd(cadd(cscale(a6.a[i], c[6]), cscale(a4.a[i], c[4])), cscale(a2.a[i], c[2])));
    for (int i = 0; i < N; ++i) v(i, i).re += c[0];
  }
  for (int i = 0; i < N * N; ++i) { p.a[i] = csub(v.a[i], u.a[i]); x.a[i] = cadd(v.a[i], u.a[i]); }
  int piv[N];
  const int sign = clu_factor(p, piv);
  for (int j = 0; j < N; ++j) clu_solve(p, piv, &x.a[j], N);
  for (int k = 0; k < squarings; ++k) {
    cmatmul(t, x, x);
    x = t;
  }
  return sign != 0;
}

// A = Q T Q^H, T upper triangular (returned in a), Q unitary.  WANTQ = false skips the accumulation (eigenvalues only).
// Returns false if the QR iteration did not converge.
template <int N, bool WANTQ>
MCM_SLA bool cschur(CMat<N>& a, CMat<N>& q) {
  if (WANTQ) cset_identity(q, 1.0);
  // -- Householder reduction to upper Hessenberg form
  for (int k = 0; k + 2 < N; ++k) {
    double tail2 = 0.0;
    for (int i = k + 2; i < N; ++i) tail2 += cabs2(a(i, k));
    if (!(tail2 > 0.0)) continue;                       // already Hessenberg in this column (or NaN)
    const double alpha = sqrt(tail2 + cabs2(a(k + 1, k)));
    cx v[N];
    for (int i = 0; i < N; ++i) v[i] = cmk(0.0, 0.0);
    for (int i = k + 1; i < N; ++i) v[i] = a(i, k);
    const double x0 = cabsv(v[k + 1]);
    const cx ph = x0 == 0.0 ? cmk(1.0, 0.0) : cscale(v[k + 1], 1.0 / x0);
    v[k + 1] = cadd(v[k + 1], cscale(ph, alpha));
    double vn2 = 0.0;
    for (int i = k + 1; i < N; ++i) vn2 += cabs2(v[i]);
    if (!(vn2 > 0.0)) continue;
    const double beta = 2.0 / vn2;
    for (int j = 0; j < N; ++j) {                       // A <- H A
      cx s = cmk(0.0, 0.0);
      for (int i = k + 1; i < N; ++i) s = cadd(s, cmul(cconj(v[i]), a(i, j)));
      s = cscale(s, beta);
      for (int i = k + 1; i < N; ++i) a(i, j) = csub(a(i, j), cmul(v[i], s));
    }
    for (int i = 0; i < N; ++i) {                       // A <- A H
      cx s = cmk(0.0, 0.0);
      for (int j = k + 1; j < N; ++j) s = cadd(s, cmul(a(i, j), v[j]));
      s = cscale(s, beta);
      for (int j = k + 1; j < N; ++j) a(i, j) = csub(a(i, j), cmul(s, cconj(v[j])));
    }
    if (WANTQ)
      for (int i = 0; i < N; ++i) {                     // Q <- Q H
        cx s = cmk(0.0, 0.0);
        for (int j = k + 1; j < N; ++j) s = cadd(s, cmul(q(i, j), v[j]));
        s = cscale(s, beta);
        for (int j = k + 1; j < N; ++j) q(i, j) = csub(q(i, j), cmul(s, cconj(v[j])));
      }
    for (int i = k + 2; i < N; ++i) a(i, k) = cmk(0.0, 0.0);
  }
  // -- shifted QR on the Hessenberg matrix, explicit steps with Givens rotations
  int nn = N - 1, its = 0;
  double cs[N];
  cx sn[N];
  while (nn > 0) {
    int l;
    for (l = nn; l > 0; --l) {
      double s = cabs1(a(l - 1, l - 1)) + cabs1(a(l, l));
      if (s == 0.0) s = 1.0;
      if (cabs1(a(l, l - 1)) <= kEps * s) { a(l, l - 1) = cmk(0.0, 0.0); break; }
    }
    if (l == nn) { --nn; its = 0; continue; }
    if (its >= 90) return false;
    cx mu;
    if (its == 10 || its == 20 || its == 40) {
      mu = cmk(a(nn, nn).re + cabs1(a(nn, nn - 1)), a(nn, nn).im);   // exceptional shift
    } else {
      // Wilkinson: the eigenvalue of the trailing 2 x 2 block closer to its last diagonal entry
      const cx aa = a(nn - 1, nn - 1), bb = a(nn - 1, nn), cc = a(nn, nn - 1), dd = a(nn, nn);
      const cx hd = cscale(csub(aa, dd), 0.5);
      const cx disc = csqrtp(cadd(cmul(hd, hd), cmul(bb, cc)));
      // dd + hd -+ disc: pick the sign that makes |hd + sgn disc| largest, mu = dd - bc / (hd + sgn disc)
      const cx p1 = cadd(hd, disc), p2 = csub(hd, disc);
      const cx den = cabs2(p1) >= cabs2(p2) ? p1 : p2;
      mu = cabs2(den) == 0.0 ? dd : csub(dd, cdiv(cmul(bb, cc), den));
    }
    ++its;
    for (int i = l; i <= nn; ++i) a(i, i) = csub(a(i, i), mu);
    for (int k = l; k < nn; ++k) {                      // R = G_{nn-1} ... G_l (H - mu I)
      const cx f = a(k, k), g = a(k + 1, k);
      const double af = cabsv(f), r = hyp(af, cabsv(g));
      double c;
      cx s;
      if (r == 0.0) { c = 1.0; s = cmk(0.0, 0.0); }
      else if (af == 0.0) { c = 0.0; s = cscale(cconj(g), 1.0 / r); }
      else { c = af / r; s = cscale(cmul(cscale(f, 1.0 / af), cconj(g)), 1.0 / r); }
      cs[k] = c; sn[k] = s;
      for (int j = k; j < N; ++j) {
        const cx t1 = a(k, j), t2 = a(k + 1, j);
        a(k, j) = cadd(cscale(t1, c), cmul(s, t2));
        a(k + 1, j) = csub(cscale(t2, c), cmul(cconj(s), t1));
      }
      a(k + 1, k) = cmk(0.0, 0.0);
    }
    for (int k = l; k < nn; ++k) {                      // H = R G_l^H ... G_{nn-1}^H + mu I
      const double c = cs[k];
      const cx s = sn[k];
      const int top = k + 1 < nn ? k + 1 : nn;
      for (int i = 0; i <= top; ++i) {
        const cx t1 = a(i, k), t2 = a(i, k + 1);
        a(i, k) = cadd(cscale(t1, c), cmul(cconj(s), t2));
        a(i, k + 1) = csub(cscale(t2, c), cmul(s, t1));
      }
      if (WANTQ)
        for (int i = 0; i < N; ++i) {
          const cx t1 = q(i, k), t2 = q(i, k + 1);
          q(i, k) = cadd(cscale(t1, c), cmul(cconj(s), t2));
          q(i, k + 1) = csub(cscale(t2, c), cmul(s, t1));
        }
    }
    for (int i = l; i <= nn; ++i) a(i, i) = cadd(a(i, i), mu);
  }
  for (int i = 1; i < N; ++i)
    for (int j = 0; j < i; ++j) a(i, j) = cmk(0.0, 0.0);
  return true;
}

// eigenvalues ordered by (re, im); false (and NaN) if the iteration did not converge
template <int N>
MCM_SLA bool ceigvals(CMat<N>& a, cx (&w)[N]) {
  CMat<N> dummy;
  const bool ok = cschur<N, false>(a, dummy);
  for (int i = 0; i < N; ++i) w[i] = ok ? a(i, i) : cmk(nan(""), nan(""));
  for (int i = 1; i < N; ++i) {
    const cx v = w[i];
    int j = i - 1;
    while (j >= 0 && (w[j].re > v.re || (w[j].re == v.re && w[j].im > v.im))) { w[j + 1] = w[j]; --j; }
    w[j + 1] = v;
  }
  return ok;
}

// a <- log(a) (principal); false if the Schur iteration failed or the matrix is singular (result non-finite)
template <int N>
MCM_SLA bool clogm(CMat<N>& a) {
  CMat<N> q;
  if (!cschur<N, true>(a, q)) {
    for (int i = 0; i < N * N; ++i) a.a[i] = cmk(nan(""), nan(""));
    return false;
  }
  CMat<N>& t = a;                                        // upper triangular
  bool ok = true;
  for (int i = 0; i < N; ++i) ok = ok && cabs2(t(i, i)) > 0.0;
  // inverse scaling: square roots of the triangular factor until it is close to the identity
  int k = 0;
  for (; k < 64; ++k) {
    double nrm = 0.0;
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int i = 0; i <= j; ++i) s += cabsv(i == j ? csub(t(i, j), cmk(1.0, 0.0)) : t(i, j));
      nrm = s > nrm || s != s ? s : nrm;
    }
    if (!(nrm > 0.264)) break;
    CMat<N> r;                                           // r r = t (Björck-Hammarling)
    for (int i = 0; i < N * N; ++i) r.a[i] = cmk(0.0, 0.0);
    for (int j = 0; j < N; ++j) {
      r(j, j) = csqrtp(t(j, j));
      for (int i = j - 1; i >= 0; --i) {
        cx s = t(i, j);
        for (int m = i + 1; m < j; ++m) s = csub(s, cmul(r(i, m), r(m, j)));
        r(i, j) = cdiv(s, cadd(r(i, i), r(j, j)));
      }
    }
    t = r;
  }
  // log(I + X) ~ sum_j w_j X (I + x_j X)^-1 at the 7 Gauss-Legendre nodes of [0, 1]
  const double gx[7] = {0.025446043828620736, 0.12923440720030277, 0.29707742431130141, 0.5, 0.70292257568869859,
                        0.87076559279969723, 0.97455395617137926};
  const double gw[7] = {0.064742483084434846, 0.13985269574463833, 0.19091502525255947, 0.20897959183673469,
                        0.19091502525255947, 0.13985269574463833, 0.064742483084434846};
  CMat<N> x, acc;
  for (int i = 0; i < N * N; ++i) { x.a[i] = t.a[i]; acc.a[i] = cmk(0.0, 0.0); }
  for (int i = 0; i < N; ++i) x(i, i).re -= 1.0;
  for (int g = 0; g < 7; ++g) {
    // solve (I + x_g X) Y = X, upper triangular, column by column
    for (int j = 0; j < N; ++j) {
      cx y[N];
      for (int i = j; i >= 0; --i) {
        cx s = x(i, j);
        for (int m = i + 1; m <= j; ++m) s = csub(s, cmul(cscale(x(i, m), gx[g]), y[m]));
        y[i] = cdiv(s, cmk(1.0 + gx[g] * x(i, i).re, gx[g] * x(i, i).im));
      }
      for (int i = 0; i <= j; ++i) acc(i, j) = cadd(acc(i, j), cscale(y[i], gw[g]));
    }
  }
  const double sc = ldexp(1.0, k);
  for (int i = 0; i < N * N; ++i) acc.a[i] = cscale(acc.a[i], sc);
  // the diagonal of log(T) is known exactly: log of the eigenvalues (principal branch), as Base recomputes it
  // (not needed for the tolerance tier; kept simple)
  // back-transformation: log(A) = Q log(T) Q^H
  CMat<N> tmp;
  cmatmul(tmp, q, acc);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      cx s = cmk(0.0, 0.0);
      for (int m = 0; m < N; ++m) s = cadd(s, cmul(tmp(i, m), cconj(q(j, m))));
      a(i, j) = s;
    }
  if (!ok) return false;
  for (int i = 0; i < N * N; ++i) ok = ok && (a.a[i].re == a.a[i].re) && (a.a[i].im == a.a[i].im);
  return ok;
}

// singular values, descending
template <int N>
MCM_SLA void csvdvals(CMat<N>& a, double (&sv)[N]) {
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < N - 1; ++p)
      for (int qq = p + 1; qq < N; ++qq) {
        double alpha = 0.0, beta = 0.0;
        cx gamma = cmk(0.0, 0.0);
        for (int i = 0; i < N; ++i) {
          alpha += cabs2(a(i, p));
          beta += cabs2(a(i, qq));
          gamma = cadd(gamma, cmul(cconj(a(i, p)), a(i, qq)));
        }
        const double g = cabsv(gamma);
        if (g == 0.0 || !(g > kEps * sqrt(alpha * beta))) continue;
        rotated = true;
        const cx ph = cscale(cconj(gamma), 1.0 / g);     // column q times conj(phase): the inner product becomes real
        const double zeta = (beta - alpha) / (2.0 * g);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
        for (int i = 0; i < N; ++i) {
          const cx aip = a(i, p), aiq = cmul(a(i, qq), ph);
          a(i, p) = csub(cscale(aip, c), cscale(aiq, s));
          a(i, qq) = cadd(cscale(aip, s), cscale(aiq, c));
        }
      }
    if (!rotated) break;
  }
  for (int j = 0; j < N; ++j) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += cabs2(a(i, j));
    sv[j] = -sqrt(s);
  }
  sort_ascending(sv);
  for (int j = 0; j < N; ++j) sv[j] = -sv[j];
}

}  // namespace sla
}  // namespace mcm
