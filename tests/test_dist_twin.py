"""Host twin of csrc/mcm_dist.cuh — quantile(d, p) for the univariate families `Particles(N, d)`, `a ⊠ d`, `a ⊞ d`
accept (src/sampling.jl:10, src/types.jl:67-90, src/particles.jl:42,50) — against scipy.stats / scipy.special and,
where scipy itself is only good to 1e-10 (the F and t upper tails), against mpmath.  Tolerance tier: 1e-12 relative.
The device runs the same source (tests/test_gpu_round2.py compares it with this twin).  CPU only."""
import ctypes as C
import os
import subprocess
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "twin", "dist_twin.cpp")
SO = os.path.join(ROOT, "tests", "twin", "_build", "libdist_twin.so")
KINDS = ["NORMAL", "UNIFORM", "EXPONENTIAL", "LOGNORMAL", "CAUCHY", "LAPLACE", "LOGISTIC", "RAYLEIGH", "WEIBULL", "PARETO", "GUMBEL",
         "TRIANGULAR", "FRECHET", "KUMARASWAMY", "GAMMA", "CHISQ", "INVERSEGAMMA", "BETA", "TDIST", "FDIST"]


def build_dist_twin():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "montecarlomeasurements.jl_b200", "csrc", f)
                    for f in ("mcm_dist.cuh", "mcm_norminv.cuh", "mcm_norminv_coeffs.inc")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO], check=True)
    return C.CDLL(SO)


def twin_quantile(lib, kind, params, p, q):
    p = np.ascontiguousarray(p, dtype=np.float64); q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty_like(p)
    a, b, c = (list(params) + [0.0, 0.0, 0.0])[:3]
    lib.twin_dist_quantile(C.c_int(KINDS.index(kind)), C.c_double(a), C.c_double(b), C.c_double(c), p.ctypes.data_as(C.c_void_p),
                           q.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(p.size))
    return out


def grid(N, idx=None):
    i = np.arange(N) if idx is None else np.asarray(idx)
    return (2 * i + 1) / (2.0 * N), (2 * (N - 1 - i) + 1) / (2.0 * N)


def cases():
    import scipy.stats as st
    out = [
        ("NORMAL", (0.5, 2.0), st.norm(0.5, 2.0), 1e-13), ("UNIFORM", (-1.0, 3.0), st.uniform(-1.0, 4.0), 1e-15),
        ("EXPONENTIAL", (2.5,), st.expon(scale=2.5), 1e-14), ("LOGNORMAL", (0.3, 0.7), st.lognorm(s=0.7, scale=np.exp(0.3)), 1e-13),
        ("CAUCHY", (1.0, 2.0), st.cauchy(1.0, 2.0), 1e-13), ("LAPLACE", (1.0, 2.0), st.laplace(1.0, 2.0), 1e-13),
        ("LOGISTIC", (1.0, 2.0), st.logistic(1.0, 2.0), 1e-13), ("RAYLEIGH", (1.5,), st.rayleigh(scale=1.5), 1e-14),
        ("WEIBULL", (1.7, 2.0), st.weibull_min(1.7, scale=2.0), 1e-14), ("PARETO", (3.0, 2.0), st.pareto(3.0, scale=2.0), 1e-14),
        ("GUMBEL", (1.0, 2.0), st.gumbel_r(1.0, 2.0), 1e-13), ("TRIANGULAR", (1.0, 4.0, 2.0), st.triang(c=1.0 / 3.0, loc=1.0, scale=3.0), 1e-12),
        ("FRECHET", (2.5, 1.5), st.invweibull(2.5, scale=1.5), 1e-14), ("INVERSEGAMMA", (3.0, 2.0), st.invgamma(3.0, scale=2.0), 1e-13),
    ]
    for a in (0.05, 0.5, 1.0, 2.0, 7.3, 100.0, 5000.0):
        out.append(("GAMMA", (a, 1.5), st.gamma(a, scale=1.5), 1e-12))
    for nu in (1.0, 3.0, 50.0):
        out.append(("CHISQ", (nu,), st.chi2(nu), 1e-12))
    for a, b in ((0.5, 0.5), (2.0, 3.0), (1.0, 1.0), (30.0, 2.0), (0.1, 5.0), (200.0, 300.0), (5.0, 0.2)):
        out.append(("BETA", (a, b), st.beta(a, b), 2e-12))
    for nu in (1.0, 2.5, 10.0, 200.0):
        out.append(("TDIST", (nu,), st.t(nu), 1e-9))            # scipy's own t quantile is only good to ~1e-10: mpmath below
    for n1, n2 in ((3.0, 7.0), (10.0, 4.0)):
        out.append(("FDIST", (n1, n2), st.f(n1, n2), 1e-9))
    return out


def test_quantiles_match_scipy_on_the_systematic_grid():
    lib = build_dist_twin()
    warnings.filterwarnings("ignore")
    p1, q1 = grid(2000)
    N2 = 10 ** 7
    p2, q2 = grid(N2, np.concatenate([np.arange(40), np.arange(N2 // 2 - 20, N2 // 2 + 20), np.arange(N2 - 40, N2)]))
    p, q = np.concatenate([p1, p2]), np.concatenate([q1, q2])
    for kind, params, d, tol in cases():
        got = twin_quantile(lib, kind, params, p, q)
        ref = np.where(p < 0.5, d.ppf(p), d.isf(q))              # the upper tail through the survival function
        scale = np.maximum(np.abs(ref), 1e-3 * abs(d.ppf(0.75) - d.ppf(0.25)))
        err = np.abs(got - ref) / scale
        assert not np.isnan(got).any() and err.max() <= tol, (kind, params, err.max(), p[np.argmax(err)])
        assert np.all(np.diff(got[:2000]) >= 0), (kind, params)                                  # monotone on the grid


def test_t_and_f_quantiles_against_mpmath():
    import mpmath as mp
    import scipy.special as sp
    import scipy.stats as st
    lib = build_dist_twin()
    mp.mp.dps = 40
    N = 2000
    for nu, i in ((1.0, 999), (2.5, 428), (10.0, 999), (200.0, 91), (3.0, 0), (7.0, 1999)):
        pp, qq = mp.mpf(2 * i + 1) / (2 * N), mp.mpf(2 * (N - 1 - i) + 1) / (2 * N)
        pm = min(pp, qq)
        f = lambda t: mp.betainc(mp.mpf(nu) / 2, mp.mpf(1) / 2, 0, mp.mpf(nu) / (nu + t * t), regularized=True) / 2 - pm
        t = abs(mp.findroot(f, abs(float(st.t(nu).ppf(float(pm)))) or 1e-6))
        truth = float(-t if pp < qq else t)
        got = twin_quantile(lib, "TDIST", (nu,), [float(pp)], [float(qq)])[0]
        assert abs(got - truth) <= 5e-13 * abs(truth), (nu, i, got, truth)
    N2 = 10 ** 7
    for (n1, n2), i in (((3.0, 7.0), N2 - 1), ((10.0, 4.0), N2 - 1), ((3.0, 7.0), 0), ((10.0, 4.0), N2 // 3)):
        pp, qq = mp.mpf(2 * i + 1) / (2 * N2), mp.mpf(2 * (N2 - 1 - i) + 1) / (2 * N2)
        g = lambda y: mp.betainc(mp.mpf(n2) / 2, mp.mpf(n1) / 2, 0, y, regularized=True) - qq
        y0 = float(sp.betaincinv(n2 / 2, n1 / 2, float(qq)))
        y = mp.findroot(g, (y0 * 0.999, min(y0 * 1.001, 1.0)), solver="illinois")
        truth = float(n2 * (1 - y) / (n1 * y))
        got = twin_quantile(lib, "FDIST", (n1, n2), [float(pp)], [float(qq)])[0]
        assert abs(got - truth) <= 1e-12 * abs(truth), (n1, n2, i, got, truth)


def test_edge_probabilities_and_parameters():
    lib = build_dist_twin()
    inf = np.inf
    for kind, params, lo, hi in (("NORMAL", (0, 1), -inf, inf), ("EXPONENTIAL", (2.0,), 0.0, inf), ("CAUCHY", (0, 1), -inf, inf),
                                 ("GAMMA", (2.0, 1.0), 0.0, inf), ("BETA", (2.0, 3.0), 0.0, 1.0), ("TDIST", (3.0,), -inf, inf),
                                 ("UNIFORM", (1.0, 2.0), 1.0, 2.0), ("PARETO", (2.0, 3.0), 3.0, inf), ("WEIBULL", (2.0, 1.0), 0.0, inf)):
        got = twin_quantile(lib, kind, params, [0.0, 1.0, 0.5, -0.1, 1.5, np.nan], [1.0, 0.0, 0.5, 1.1, -0.5, np.nan])
        assert got[0] == lo and got[1] == hi and np.isfinite(got[2]) and np.isnan(got[3:]).all(), (kind, got)
    assert np.isnan(twin_quantile(lib, "GAMMA", (-1.0, 1.0), [0.3], [0.7])[0])
    assert np.isnan(twin_quantile(lib, "BETA", (0.0, 1.0), [0.3], [0.7])[0])
