"""Pin the CPU oracle: against the reference's own published known-answer values, against mpmath
(the arbiter of the ulp tier) and against numpy for the exact / order-statistic tiers.  CPU only."""
import numpy as np
import pytest

from conftest import bits_equal, ulp_err


def test_readme_pstd_pi_pm_01(orc, golden):
    # README.md:31-32: pstd(pi ± 0.1) = 0.09999231528930486  (N = 2000 systematic sample)
    a = orc.systematic_normal(2000, np.pi, 0.1)
    want = golden["known"]["pstd_pi_pm_0.1"]["value"]
    assert abs(orc.reduce("STD", a) - want) <= 1e-12 * want
    # permutation invariance (sampling.jl:11 permutes; statistics do not care)
    perm = orc.randperm(2000, seed=3)
    b = orc.systematic_normal(2000, np.pi, 0.1, perm=perm)
    assert bits_equal(np.sort(a), np.sort(b))
    assert abs(orc.reduce("STD", b) - want) <= 1e-12 * want


def test_docs_static_particles_100(orc, golden):
    # docs/src/index.md:55-62: p = StaticParticles(100): pstd, pvar, pmean ~ -6.66e-17
    p = orc.systematic_normal(100, 0.0, 1.0)
    k = golden["known"]
    assert abs(orc.reduce("STD", p) - k["static100_pstd"]["value"]) <= 1e-12
    assert abs(orc.reduce("VAR", p) - k["static100_pvar"]["value"]) <= 1e-12
    assert abs(orc.reduce("MEAN", p)) < 1e-15
    # docs/src/index.md:73-74: Normal(f(p)), f(x) = 2x + 10
    y = orc.binary("ADD", orc.binary("MUL", p, 2.0), 10.0)
    assert abs(orc.reduce("MEAN", y) - k["normal_fit_2x_plus_10"]["mu"]) <= 1e-13
    assert abs(orc.reduce("STD", y) - k["normal_fit_2x_plus_10"]["sigma"]) <= 1e-12
    # README.md:28-30: 2 ∓ 0.1 prints 2.0 ± 0.0999
    q = orc.systematic_normal(100, 2.0, 0.1)
    assert round(orc.reduce("STD", q), 4) == k["two_mp_0.1_pstd_printed"]["value"]


def test_docs_prob_lt_1(orc, golden):
    # docs/src/index.md:346-354: @prob Particles() < 1 == 0.8415  (1683 of 2000)
    p = orc.systematic_normal(2000, 0.0, 1.0)
    assert orc.count("LT", p, 1.0) == 1683
    assert orc.count("LT", p, 1.0) / 2000 == golden["known"]["prob_lt_1"]["value"]


def test_norminvcdf_against_mpmath(orc, golden):
    got = orc.norminvcdf(golden["norminv_p"])
    u = ulp_err(got, golden["norminv_y"])
    assert u.max() <= 1.0, u.max()
    assert np.isneginf(orc.norminvcdf(0.0)[0]) and np.isposinf(orc.norminvcdf(1.0)[0])
    assert np.isnan(orc.norminvcdf(1.5)[0])


@pytest.mark.parametrize("name,fn", [("SIN", "sin"), ("COS", "cos"), ("EXP", "exp"), ("LOG", "log"), ("TAN", "tan"),
                                     ("ASIN", "asin"), ("ACOS", "acos"), ("ATAN", "atan"), ("SINH", "sinh"),
                                     ("COSH", "cosh"), ("TANH", "tanh"), ("EXPM1", "expm1"), ("LOG1P", "log1p"),
                                     ("SQRT", "sqrt")])
def test_oracle_transcendentals_within_1ulp_of_truth(orc, golden, name, fn):
    x = golden[fn + "_x"]
    u = ulp_err(orc.unary(name, x), golden[fn + "_y"])
    lim = 2.0 if name in ("TANH", "SINH", "COSH", "TAN") else 1.0   # glibc's documented bounds
    assert u.max() <= lim, (name, u.max())


def test_oracle_pow_truth(orc, golden):
    got = orc.binary("POW", golden["pow_x"], golden["pow_e"])
    assert ulp_err(got, golden["pow_y"]).max() <= 1.0


def test_exact_tier_matches_ieee(orc):
    rng = np.random.default_rng(1)
    for dt in (np.float64, np.float32):
        x = rng.normal(size=1000).astype(dt); y = rng.normal(size=1000).astype(dt)
        assert bits_equal(orc.binary("ADD", x, y), x + y)
        assert bits_equal(orc.binary("SUB", x, y), x - y)
        assert bits_equal(orc.binary("MUL", x, y), x * y)
        assert bits_equal(orc.binary("DIV", x, y), x / y)
        assert bits_equal(orc.binary("SUB", 2.5, y), dt(2.5) - y)         # f(a::Real, p) form
        assert bits_equal(orc.unary("SQR", x), x * x)                      # literal_pow 2
        assert bits_equal(orc.unary("CUBE", x), x * x * x)
        assert bits_equal(orc.powi(x, 3), x * x * x)
        assert bits_equal(orc.powi(x, -1), dt(1) / x)
        assert bits_equal(orc.unary("SQRT", np.abs(x)), np.sqrt(np.abs(x)))
        assert bits_equal(orc.unary("ROUND", x * 7), np.rint(x * 7))       # RoundNearest, ties to even
        assert bits_equal(orc.ternary("CLAMP", x, -0.5, 0.5), np.clip(x, dt(-0.5), dt(0.5)))
        assert bits_equal(orc.ternary("MUL_ADD", x, y, 0.25), x * y + dt(0.25))
        assert bits_equal(orc.binary("REM", x, y), np.fmod(x, y))
        assert bits_equal(orc.binary("MOD", x, y), np.mod(x, y))           # floored modulo == Base.mod for floats


def test_julia_minmax_sign_semantics(orc):
    x = np.array([0.0, -0.0, 1.0, np.nan, 2.0, -np.inf])
    y = np.array([-0.0, 0.0, np.nan, 1.0, 2.0, np.inf])
    mx = orc.binary("MAX", x, y); mn = orc.binary("MIN", x, y)
    assert not np.signbit(mx[0]) and not np.signbit(mx[1])        # max(0.0,-0.0) = 0.0
    assert np.signbit(mn[0]) and np.signbit(mn[1])                 # min(0.0,-0.0) = -0.0
    assert np.isnan(mx[2]) and np.isnan(mx[3]) and np.isnan(mn[2]) and np.isnan(mn[3])  # NaN propagates
    s = orc.unary("SIGN", np.array([-2.0, 0.0, -0.0, 3.0, np.nan]))
    assert s[0] == -1 and s[3] == 1 and s[1] == 0 and np.signbit(s[2]) and np.isnan(s[4])
    assert orc.binary("MOD1", np.array([4.0, 5.0, -1.0]), 4.0).tolist() == [4.0, 1.0, 3.0]
    sd = orc.unary("SIND", np.array([30.0, 180.0, -180.0, 360.0, 90.0]))
    assert abs(sd[0] - 0.5) < 1e-16 and sd[1] == 0 and np.signbit(sd[2]) and sd[3] == 0 and sd[4] == 1.0
    assert orc.unary("COSD", np.array([90.0, 60.0, 270.0]))[0] == 0.0


def test_statistics_tolerance_tier(orc):
    rng = np.random.default_rng(2)
    x = 1e6 + rng.normal(size=20001)
    m = orc.moments(x)
    assert m[0] == x.size
    assert abs(m[1] - np.mean(x)) <= 1e-12 * abs(np.mean(x))
    assert abs(m[2] / (x.size - 1) - np.var(x - 1e6, ddof=1)) <= 1e-10
    assert m[3] == x.min() and m[4] == x.max()
    assert orc.reduce("STD", np.array([3.0])) == 0.0 and orc.reduce("VAR", np.array([3.0])) == 0.0  # N == 1 -> zero(T)
    X = rng.normal(size=(6, 500)); X[3] = 2 * X[1] + 1
    C = orc.cov(X)
    assert np.allclose(C, np.cov(X), rtol=1e-12, atol=1e-14)
    assert abs(C[3, 1] - 2 * C[1, 1]) < 1e-12


def test_quantile_type7_and_sort_order(orc):
    rng = np.random.default_rng(3)
    x = rng.normal(size=1001)
    q = np.array([0.0, 0.025, 0.25, 0.5, 0.975, 1.0, 1 / 3])
    assert np.allclose(orc.quantile(x, q), np.quantile(x, q), rtol=0, atol=1e-14)
    # Float32 data with Float64 probabilities promotes like Julia
    xf = x.astype(np.float32)
    assert np.allclose(orc.quantile(xf, q), np.quantile(xf.astype(np.float64), q), atol=1e-6)
    z = np.array([0.0, -0.0, 1.0, -1.0, np.nan, -0.0, 0.0])
    s = orc.sort(z)
    assert np.isnan(s[-1]) and s[0] == -1.0 and np.signbit(s[1]) and np.signbit(s[2]) and not np.signbit(s[3])
    assert orc.sortperm(z).tolist() == [3, 1, 5, 0, 6, 2, 4]   # stable, -0.0 < 0.0, NaN last
    with pytest.raises(ValueError):
        orc.quantile(z, [0.5])                                   # NaN -> ArgumentError in Statistics


def test_resampling_restatement(orc):
    rng = np.random.default_rng(4)
    idx = orc.rand_indices(5000, 777, seed=9)
    assert idx.min() >= 0 and idx.max() < 777
    assert np.array_equal(idx, orc.rand_indices(5000, 777, seed=9)) and not np.array_equal(idx, orc.rand_indices(5000, 777, seed=10))
    src = rng.normal(size=777)
    assert bits_equal(orc.gather(src, idx), src[idx])
    p = orc.randperm(1000, seed=5)
    assert np.array_equal(np.sort(p), np.arange(1000))
    # systematic resampling: equal weights -> identity; the dead reference loop (resampling.jl:45-67) in floats
    n = 64
    assert np.array_equal(orc.resample_systematic(np.ones(n), 0.5), np.arange(n))
    w = rng.random(n) + 0.01
    got = orc.resample_systematic(w, 0.3)
    S = w.sum(); s = 0.3 * S / n; b = 0; binv = w[0]; ref = []
    for i in range(n):
        while not (s < binv):
            b += 1; binv += w[b]
        ref.append(b); s += S / n
    assert got.tolist() == ref
    assert orc.resample_systematic(np.array([0.0, 0.0, 5.0, 0.0]), 0.9).tolist() == [2, 2, 2, 2]


def test_oracle_decision_restatements(orc):
    """oracle.union / intersect / with_nominal against the reference's own assertions (test/runtests.jl:525-527,
    :930-937) and Julia's argmin rules (first minimum; NaN wins)."""
    o = orc
    x = o.systematic_normal(2000, 1.0, 0.1)
    assert np.array_equal(o.intersect(x, x), o.union(x, x))
    assert o.intersect(x, 1 + x).size < 2 * x.size and o.union(x, 1 + x).size == 2 * x.size
    pn = o.with_nominal(x, 0.0)
    assert pn[0] == 0.0 and x[0] != 0.0 and np.sum(pn != x) <= 2
    assert o.julia_argmin(np.array([2.0, 1.0, 1.0])) == 1
    assert o.julia_argmin(np.array([2.0, np.nan, 1.0, np.nan])) == 1
