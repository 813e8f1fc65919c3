"""GPU parity: the CUDA path (through the host mirror -> ctypes -> C ABI of libmcm_b200.so) against the
CPU oracle and the committed golden fixtures, on seeded inputs the oracle finishes in seconds.

Tolerances (BASELINE.json north_star): + - * / and index outputs bit-exact; transcendentals within 2 ulp;
Float64 reductions within 1e-12 relative (Float32: 1e-5)."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal, ulp_err

pytestmark = pytest.mark.gpu

N = 40003  # odd: exercises the 128-bit tail path
DTYPES = [np.float64, np.float32]


def rand_inputs(dt, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(0.3, 1.0, N).astype(dt)
    x[:8] = [0.0, -0.0, 1.0, -1.0, 2.5, -2.5, 1e-3, 30.0]
    return x, rng.normal(-0.2, 2.0, N).astype(dt)


EXACT_UNARY = ["PLUS", "NEG", "ZERO", "SIGN", "ABS", "FLOAT", "ROUND", "FLOOR", "CEIL", "TRUNC", "SQR", "CUBE",
               "RAD2DEG", "DEG2RAD"]
ULP_UNARY = {  # name -> (domain map, ulp limit vs the oracle = the correctly rounded value, oracle/mcm_oracle.c u64/u32)
    # north_star: transcendentals within 2 ulp.  Measured on a B200 over 2e5 points (scripts/probe_ulps.py,
    # profiles/r2_ulps.txt): every Float64 function <= 2 from the truth, hand-written sin/cos/exp <= 1.
    "EXP": (lambda x: x, 2), "EXP2": (lambda x: x, 2), "EXP10": (lambda x: x, 2), "EXPM1": (lambda x: x, 2),
    "LOG": (lambda x: np.abs(x) + 0.01, 2), "LOG10": (lambda x: np.abs(x) + 0.01, 2),
    "LOG2": (lambda x: np.abs(x) + 0.01, 2), "LOG1P": (lambda x: np.abs(x), 2),
    "SIN": (lambda x: x * 3, 2), "COS": (lambda x: x * 3, 2), "TAN": (lambda x: x, 2),
    "SINH": (lambda x: x * 2, 2), "COSH": (lambda x: x * 2, 2), "TANH": (lambda x: x, 2),
    "ASIN": (lambda x: np.clip(x, -0.99, 0.99), 2), "ACOS": (lambda x: np.clip(x, -0.99, 0.99), 2),
    "ATAN": (lambda x: x * 5, 2), "ASIND": (lambda x: np.clip(x, -0.99, 0.99), 2),
    "ACOSD": (lambda x: np.clip(x, -0.99, 0.99), 2), "ATAND": (lambda x: x * 5, 2),
    "ASINH": (lambda x: x * 5, 2), "ACOSH": (lambda x: np.abs(x) + 1.01, 2), "ATANH": (lambda x: np.clip(x, -0.99, 0.99), 2),
}


@pytest.mark.parametrize("dt", DTYPES)
def test_exact_tier_unary_bit_exact(mcm, orc, dt):
    x, _ = rand_inputs(dt)
    p = mcm.Particles(x * dt(7.3))
    for name in EXACT_UNARY:
        got = mcm.particles._unary(name, p).to_numpy()
        assert bits_equal(got, orc.unary(name, x * dt(7.3))), name
    xp = np.abs(x) + dt(0.01)
    pp = mcm.Particles(xp)
    assert bits_equal(mcm.sqrt(pp).to_numpy(), orc.unary("SQRT", xp))      # IEEE correctly rounded
    assert bits_equal(mcm.inv(pp).to_numpy(), orc.unary("INV", xp))


@pytest.mark.parametrize("dt", DTYPES)
def test_ulp_tier_unary_vs_oracle(mcm, orc, dt):
    x, _ = rand_inputs(dt, 1)
    for name, (dom, lim) in ULP_UNARY.items():
        xi = np.ascontiguousarray(dom(x).astype(dt))
        got = mcm.particles._unary(name, mcm.Particles(xi)).to_numpy()
        u = ulp_err(got, orc.unary(name, xi))
        assert u.max() <= lim, (name, np.dtype(dt).name, u.max())


def test_float64_transcendentals_vs_mpmath_truth(mcm, golden):
    """The hot SLEEF-class primitives against correctly rounded truth (tests/golden, mpmath)."""
    limits = {"sin": 2, "cos": 2, "exp": 1, "log": 1, "tan": 2, "asin": 2, "acos": 2, "atan": 2, "sinh": 2, "cosh": 2,
              "tanh": 2, "expm1": 2, "log1p": 2, "sqrt": 0}
    worst = {}
    for fn, lim in limits.items():
        x = golden[fn + "_x"]
        got = getattr(mcm, fn)(mcm.Particles(x)).to_numpy()
        u = ulp_err(got, golden[fn + "_y"])
        worst[fn] = u.max()
        assert u.max() <= lim, (fn, u.max())
    got = (mcm.Particles(golden["pow_x"]) ** mcm.Particles(golden["pow_e"])).to_numpy()
    assert ulp_err(got, golden["pow_y"]).max() <= 2
    half = golden["pow_e"] == 1.5
    got15 = (mcm.Particles(golden["pow_x"][half]) ** 1.5).to_numpy()       # scalar exponent form f(p, a::Real)
    assert ulp_err(got15, golden["pow_y"][half]).max() <= 2


@pytest.mark.parametrize("dt", DTYPES)
def test_degree_trig_exact_points_and_accuracy(mcm, orc, dt):
    pts = np.array([0, 30, 45, 60, 90, 120, 180, 270, 360, -30, -90, -180, 720, 1e6 + 30], dtype=dt)
    s = mcm.sind(mcm.Particles(pts)).to_numpy(); c = mcm.cosd(mcm.Particles(pts)).to_numpy()
    assert s[0] == 0 and s[4] == 1 and s[6] == 0 and s[8] == 0 and s[10] == -1 and np.signbit(s[11])
    assert c[4] == 0 and c[7] == 0 and c[0] == 1 and c[6] == -1
    assert abs(float(s[1]) - 0.5) <= np.finfo(dt).eps and abs(float(c[3]) - 0.5) <= np.finfo(dt).eps
    x, _ = rand_inputs(dt, 2)
    xd = (x * 200).astype(dt)
    for name in ("SIND", "COSD"):
        got = mcm.particles._unary(name, mcm.Particles(xd)).to_numpy().astype(np.float64)
        ref = orc.unary(name, xd).astype(np.float64)
        assert np.max(np.abs(got - ref)) <= 2 * np.finfo(dt).eps, name


@pytest.mark.parametrize("dt", DTYPES)
def test_binary_forms_and_exactness(mcm, orc, dt):
    x, y = rand_inputs(dt, 3)
    px, py = mcm.Particles(x), mcm.Particles(y)
    # f(p1, p2): + - * / bit-exact (built with -fmad=false)
    for name, got in (("ADD", px + py), ("SUB", px - py), ("MUL", px * py), ("DIV", px / py),
                      ("MAX", mcm.maximum(px, py)), ("MIN", mcm.minimum(px, py)), ("REM", mcm.rem(px, py)),
                      ("MOD", mcm.mod(px, py)), ("MOD1", mcm.mod1(px, py)), ("FDIV", mcm.div(px, py)),
                      ("ADD_SUM", mcm.add_sum(px, py))):
        assert bits_equal(got.to_numpy(), orc.binary(name, x, y)), name
    # f(p, a::Real) and f(a::Real, p)
    assert bits_equal((px * 2.5).to_numpy(), orc.binary("MUL", x, 2.5))
    assert bits_equal((2.5 - px).to_numpy(), orc.binary("SUB", 2.5, x))
    assert bits_equal((1.0 / py).to_numpy(), orc.binary("DIV", 1.0, y))
    assert bits_equal((px + 1).to_numpy(), orc.binary("ADD", x, 1.0))
    # three-argument forms p + p + 1, clamp(p, lo, hi) (register_primitive.jl:44-47)
    assert bits_equal((px + py + 1).to_numpy(), orc.binary("ADD", orc.binary("ADD", x, y), 1.0))
    assert bits_equal(mcm.clamp(px, -0.5, 0.7).to_numpy(), orc.ternary("CLAMP", x, -0.5, 0.7))
    assert bits_equal(mcm.muladd(px, py, 0.25).to_numpy(), orc.ternary("MULADD", x, y, 0.25))
    for name, got in (("ATAN2", mcm.atan(px, py)), ("ATAND2", mcm.atand(px, py)), ("HYPOT", mcm.hypot(px, py))):
        assert ulp_err(got.to_numpy(), orc.binary(name, x, y)).max() <= 2, name
    xp = np.abs(x) + dt(0.01)
    got = (mcm.Particles(xp) ** py).to_numpy()
    assert ulp_err(got, orc.binary("POW", xp, y)).max() <= 2
    # Julia max/min semantics: NaN propagates, -0.0 < 0.0
    a = mcm.Particles(np.array([0.0, -0.0, 1.0, np.nan], dtype=dt)); b = mcm.Particles(np.array([-0.0, 0.0, np.nan, 1.0], dtype=dt))
    mx, mn = mcm.maximum(a, b).to_numpy(), mcm.minimum(a, b).to_numpy()
    assert not np.signbit(mx[0]) and not np.signbit(mx[1]) and np.signbit(mn[0]) and np.signbit(mn[1])
    assert np.isnan(mx[2:]).all() and np.isnan(mn[2:]).all()


@pytest.mark.parametrize("dt", DTYPES)
def test_integer_powers(mcm, orc, dt):
    x, _ = rand_inputs(dt, 4)
    xp = np.abs(x) + dt(0.05)
    p = mcm.Particles(xp)
    for k in (-2, -1, 0, 1, 2, 3):                     # literal_pow: bit-exact
        assert bits_equal((p ** k).to_numpy(), orc.powi(xp, k)), k
    for k in (4, 5, 7, 10, -3, -7, 17):                # compensated power: within 1 ulp of the exact power
        assert ulp_err((p ** k).to_numpy(), orc.powi(xp, k)).max() <= 1, k
    assert bits_equal((mcm.Particles(x) ** 2).to_numpy(), x * x)


@pytest.mark.parametrize("dt", DTYPES)
def test_mixed_eltype_promotes(mcm, dt):
    a = mcm.Particles(np.arange(8, dtype=np.float32)); b = mcm.Particles(np.arange(8, dtype=np.float64))
    assert (a + b).np_dtype == np.float64 and (a * 2.0).np_dtype == np.float32
    assert np.array_equal((a + b).to_numpy(), 2.0 * np.arange(8))
    with pytest.raises(mcm.MethodError):               # N is a type parameter: mismatch is a MethodError
        mcm.Particles(np.zeros(4, dtype=dt)) + mcm.Particles(np.zeros(5, dtype=dt))
    with pytest.raises(mcm.MethodError):
        mcm.ratdiv(a, a)                                # `//` has no float method


@pytest.mark.parametrize("dt", DTYPES)
def test_fused_affine_poly_lincomb(mcm, orc, dt):
    x, y = rand_inputs(dt, 5)
    px, py = mcm.Particles(x), mcm.Particles(y)
    assert bits_equal(mcm.affine(px, 1.25, -3.0).to_numpy(), orc.binary("ADD", orc.binary("MUL", x, 1.25), -3.0))
    coeffs = [0.5, -1.0, 0.25, 2.0, -0.125, 0.3, 0.01, -0.02, 0.003]        # degree 8
    acc = np.full_like(x, dt(coeffs[-1]))
    accf = np.full_like(x, dt(coeffs[-1]))
    for c in coeffs[-2::-1]:
        acc = orc.ternary("MUL_ADD", acc, x, c); accf = orc.ternary("MULADD", accf, x, c)
    assert bits_equal(mcm.evalpoly(px, coeffs, fused=False).to_numpy(), acc)
    assert bits_equal(mcm.evalpoly(px, coeffs, fused=True).to_numpy(), accf)
    z = (np.abs(x) + 1).astype(dt); pz = mcm.Particles(z)
    ref = orc.binary("ADD", orc.binary("ADD", orc.binary("MUL", x, 0.5), orc.binary("MUL", y, 2.0)), orc.binary("MUL", z, -1.5))
    assert bits_equal(mcm.lincomb([0.5, 2.0, -1.5], [px, py, pz]).to_numpy(), ref)


@pytest.mark.parametrize("dt", DTYPES)
def test_chain10_fused_vs_oracle_and_engines_agree(mcm, orc, dt):
    from mcm_b200 import chains
    L = mcm._lib
    rng = np.random.default_rng(6)
    a = (np.pi + 0.1 * rng.normal(size=N)).astype(dt)
    pa = mcm.Particles(a)
    f = mcm.fused(chains.chain10)
    y, st = f(pa, stats=True)
    assert L.lib().mcm_chain_last_engine() in (1, 3)                # a straight-line kernel: registered or generated
    got = y.to_numpy()
    ref = orc.chain10(a)
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= orc.chain10_bound(a))
    # fused statistics == statistics of the materialised result
    m = orc.moments(got)
    tol = 1e-12 if dt == np.float64 else 1e-5
    assert st.n == N and abs(st.mean - m[1]) <= tol * abs(m[1]) and abs(st.m2 - m[2]) <= tol * m[2]
    assert st.min == m[3] and st.max == m[4]
    # statistics-only launch (y never written) gives the same numbers
    none, st2 = f(pa, stats=True, materialize=False)
    assert none is None and st2.mean == st.mean and st2.m2 == st.m2
    # the interpreter engine computes bit-identical values
    L.lib().mcm_chain_force_interpreter(1)
    try:
        y2 = f(pa)
        assert L.lib().mcm_chain_last_engine() == 2
    finally:
        L.lib().mcm_chain_force_interpreter(0)
    assert bits_equal(y2.to_numpy(), got)
    # so do the registered (ahead-of-time) kernel and the kernel generated from the tape at run time
    for eng in (1, 3):
        L.check(L.lib().mcm_chain_force_engine(eng))
        try:
            ye, ste = f(pa, stats=True)
            assert L.lib().mcm_chain_last_engine() == eng
        finally:
            L.check(L.lib().mcm_chain_force_engine(0))
        assert bits_equal(ye.to_numpy(), got)
        assert ste.n == N and abs(ste.mean - st.mean) <= 1e-13 * abs(st.mean) and abs(ste.m2 - st.m2) <= 1e-12 * st.m2
    # unfused evaluation (one launch per primitive, like the reference) is bit-identical too
    y3 = mcm.sqrt(abs(mcm.sin(pa) * pa ** 2 + mcm.exp(-pa))) / pa ** 1.5
    assert bits_equal(y3.to_numpy(), got)


def test_unregistered_chain_runs_on_interpreter(mcm, orc):
    rng = np.random.default_rng(7)
    a = rng.normal(size=N); b = np.abs(rng.normal(size=N)) + 0.5
    f = mcm.fused(lambda p, q: (mcm.log(q) * p - mcm.cos(p / q) ** 3, mcm.clamp(p, -1.0, 1.0) + 2.0 * q))
    y0, y1 = f(mcm.Particles(a), mcm.Particles(b))
    assert mcm._lib.lib().mcm_chain_last_engine() == 2
    r0 = orc.binary("SUB", orc.binary("MUL", orc.unary("LOG", b), a), orc.unary("CUBE", orc.unary("COS", orc.binary("DIV", a, b))))
    r1 = orc.binary("ADD", orc.ternary("CLAMP", a, -1.0, 1.0), orc.binary("MUL", b, 2.0))
    assert np.all(np.abs(y0.to_numpy() - r0) <= 8 * np.finfo(np.float64).eps * (np.abs(r0) + np.abs(np.log(b) * a)))
    assert bits_equal(y1.to_numpy(), r1)


def test_register_primitive(mcm, orc):
    """runtests.jl:972-975: after register_primitive(testf), testf(p, p) == p + p elementwise."""
    testf = mcm.register_primitive(lambda x, y: x + y, "testf")
    p = mcm.Particles(np.random.default_rng(8).normal(size=2000))
    assert bits_equal(testf(p, p).to_numpy(), (p + p).to_numpy())
    g = mcm.register_primitive(lambda x, s: mcm.exp(x * s) - 1.0, "g")
    e = orc.unary("EXP", orc.binary("MUL", p.to_numpy(), 0.5))
    assert np.all(np.abs(g(p, 0.5).to_numpy() - orc.binary("SUB", e, 1.0)) <= 2 * np.finfo(np.float64).eps * e)
    assert testf(2.0, 3.0) == 5.0                                      # plain numbers pass through


@pytest.mark.parametrize("dt", DTYPES)
def test_pendulum_rk4_in_register_loop(mcm, orc, dt):
    from mcm_b200 import chains, tape
    n, steps = 5001, 200
    rng = np.random.default_rng(9)
    th = np.zeros(n, dtype=dt); om = (np.pi / 3 + 0.02 * rng.normal(size=n)).astype(dt)
    g = (9.79 + 0.02 * rng.normal(size=n)).astype(dt); Lp = (1 + 0.01 * rng.normal(size=n)).astype(dt)
    c = (0.1 + 0.01 * rng.normal(size=n)).astype(dt)
    t = chains.pendulum_rk4_tape(0.01)
    ins = [mcm.Particles(v) for v in (th, om, g, Lp, c)]
    o1, o2 = mcm.particles._empty(n, ins[0].dtype), mcm.particles._empty(n, ins[0].dtype)
    tape.run_tape(t, [p.ptr for p in ins], n, ins[0].dtype, [o1.ptr, o2.ptr], reps=steps)
    assert mcm._lib.lib().mcm_chain_last_engine() in (1, 3)
    # the three engines agree bit for bit on the loop tape
    ref_bits = (o1.to_numpy(), o2.to_numpy())
    for eng in (1, 2, 3):
        mcm._lib.check(mcm._lib.lib().mcm_chain_force_engine(eng))
        try:
            q1, q2 = mcm.particles._empty(n, ins[0].dtype), mcm.particles._empty(n, ins[0].dtype)
            tape.run_tape(t, [p.ptr for p in ins], n, ins[0].dtype, [q1.ptr, q2.ptr], reps=steps)
            assert mcm._lib.lib().mcm_chain_last_engine() == eng
        finally:
            mcm._lib.check(mcm._lib.lib().mcm_chain_force_engine(0))
        assert bits_equal(q1.to_numpy(), ref_bits[0]) and bits_equal(q2.to_numpy(), ref_bits[1]), eng
    rth, rom = orc.pendulum_rk4(th.astype(np.float64), om.astype(np.float64), g.astype(np.float64),
                                Lp.astype(np.float64), c.astype(np.float64), 0.01, steps)
    tol = 1e-11 if dt == np.float64 else 2e-4
    assert np.max(np.abs(o1.to_numpy() - rth)) <= tol and np.max(np.abs(o2.to_numpy() - rom)) <= tol


@pytest.mark.parametrize("dt", DTYPES)
def test_statistics(mcm, orc, dt):
    x, _ = rand_inputs(dt, 10)
    p = mcm.Particles(x)
    m = orc.moments(x)
    tol = 1e-12 if dt == np.float64 else 1e-5
    assert abs(mcm.pmean(p) - m[1]) <= tol * max(abs(m[1]), 1e-300) + 1e-16
    assert abs(mcm.pvar(p) - m[2] / (N - 1)) <= tol * m[2] / (N - 1)
    assert abs(mcm.pstd(p) - orc.reduce("STD", x)) <= tol * orc.reduce("STD", x)
    assert abs(mcm.pvar(p, corrected=False) - m[2] / N) <= tol * m[2] / N
    assert mcm.pminimum(p) == m[3] and mcm.pmaximum(p) == m[4] and mcm.pextrema(p) == (m[3], m[4])
    assert abs(mcm.psum(p) - m[5]) <= tol * np.abs(x.astype(np.float64)).sum()
    assert abs(mcm.pstd(p, mean=0.25) - np.sqrt(((x.astype(np.float64) - 0.25) ** 2).sum() / (N - 1))) <= 1e-5
    assert abs(mcm.meanstd(p) - mcm.pstd(p) / np.sqrt(N)) < 1e-15
    # N == 1 -> zero(T) (particles.jl:309)
    one = mcm.Particles(np.array([3.0], dtype=dt))
    assert mcm.pstd(one) == 0.0 and mcm.pvar(one) == 0.0 and mcm.pmean(one) == 3.0
    # pmean(::Number) passthrough and vector map (particles.jl:322,475)
    assert mcm.pmean(2.5) == 2.5 and mcm.pmean([p, one]) == [mcm.pmean(p), 3.0]
    # NaN propagates through mean and extrema
    xn = x.copy(); xn[17] = np.nan
    q = mcm.Particles(xn)
    assert np.isnan(mcm.pmean(q)) and np.isnan(mcm.pminimum(q)) and np.isnan(mcm.pmaximum(q))


def test_moments_no_cancellation_at_large_offset(mcm, orc):
    rng = np.random.default_rng(11)
    for off in (1e9, -1e12, 1e15):
        x = off + rng.normal(size=N) * (1 + abs(off) * 1e-12)
        got = mcm.pvar(mcm.Particles(x))
        want = orc.moments(x)[2] / (N - 1)
        assert abs(got - want) <= 1e-12 * want, off
    # an outlier as the very first element must not poison the shifted accumulation
    x = rng.normal(size=N); x[0] = 1e10
    assert abs(mcm.pvar(mcm.Particles(x)) - orc.moments(x)[2] / (N - 1)) <= 1e-12 * orc.moments(x)[2] / (N - 1)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 256, 257, 2047, 2049, 303105])
def test_ragged_sizes(mcm, orc, n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n)
    p = mcm.Particles(x)
    assert bits_equal((p * p + 1.0).to_numpy(), x * x + 1.0)
    assert bits_equal(mcm.psort(p).to_numpy(), orc.sort(x))
    assert abs(mcm.pmean(p) - x.mean()) <= 1e-12 * max(abs(x.mean()), 1e-3)
    from mcm_b200 import chains
    y, st = mcm.fused(chains.readme_expr)(p, stats=True)
    assert np.all(np.abs(y.to_numpy() - orc.readme_expr(x)) <= 8e-16 * (np.abs(np.sin(x) * x * x) + np.exp(-x)))
    assert st.n == n
    assert mcm.pquantile(p, 0.5) == orc.quantile(x, 0.5)[0]


@pytest.mark.parametrize("dt", DTYPES)
def test_radix_select_edge_cases(mcm, orc, dt):
    """MSD radix select must return exact order statistics for degenerate key distributions."""
    rng = np.random.default_rng(31)
    q = np.array([0.0, 0.01, 0.25, 0.5, 0.75, 0.99, 1.0, 1 / 3])
    cases = {
        "constant": np.full(100_001, 2.5), "two values": np.repeat([1.0, -3.0], 60_000),
        "small ints": rng.integers(0, 50, 300_000).astype(np.float64),
        "wide range": np.exp(rng.uniform(-80, 80, 200_000)) * rng.choice([-1.0, 1.0], 200_000),
        "with inf": np.concatenate([rng.normal(size=5000), [np.inf, -np.inf, np.inf]]),
        "signed zeros": np.concatenate([np.zeros(3000), -np.zeros(3001)]),
        "n=1": np.array([4.0]), "n=2": np.array([4.0, -1.0]), "n=4097": rng.normal(size=4097),
        "heavy ties": np.round(rng.normal(size=500_000), 1),
    }
    for name, x in cases.items():
        x = x.astype(dt)
        got = mcm.pquantile(mcm.Particles(x), q)
        want = orc.quantile(x, q)
        assert np.array_equal(got, want) or (np.isnan(got) == np.isnan(want)).all() and np.array_equal(np.nan_to_num(got), np.nan_to_num(want)), name


def test_empty_and_unaligned_views(mcm):
    L = mcm._lib
    e = mcm.Particles(np.zeros(0))
    assert (e + 1.0).N == 0
    with pytest.raises(mcm.ArgumentError):
        mcm.pmean(e)                                   # mean of an empty collection
    # an odd-element offset view (8-byte aligned only) goes through the scalar kernel and is still exact
    x = np.random.default_rng(12).normal(size=1025)
    p = mcm.Particles(x)
    out = mcm.particles._empty(1024, p.dtype)
    L.check(L.lib().mcm_unary(L.ENUMS["MCM_U_SQR"], p.dtype, C.c_void_p(p.ptr + 8), C.c_void_p(out.ptr), 1024))
    assert bits_equal(out.to_numpy(), x[1:] * x[1:])
    L.check(L.lib().mcm_binary(L.ENUMS["MCM_B_ADD"], p.dtype, C.c_void_p(p.ptr + 8), 0.0, C.c_void_p(p.ptr), 0.0,
                               C.c_void_p(out.ptr), 1024))
    assert bits_equal(out.to_numpy(), x[1:] + x[:-1])
    m = np.zeros(6)
    L.check(L.lib().mcm_moments(p.dtype, C.c_void_p(p.ptr + 8), 1024, m.ctypes.data_as(L._dp)))
    assert abs(m[1] - x[1:].mean()) < 1e-14


@pytest.mark.parametrize("dt", DTYPES)
def test_sort_sortperm_quantile_bit_exact(mcm, orc, dt):
    x, _ = rand_inputs(dt, 13)
    x[100:110] = 0.0; x[110:120] = -0.0; x[200:260] = x[300]          # ties, signed zeros
    p = mcm.Particles(x)
    assert bits_equal(mcm.psort(p).to_numpy(), orc.sort(x))
    assert np.array_equal(mcm.psortperm(p).to_numpy(), orc.sortperm(x))    # stable, 0-based
    q = np.array([0.0, 0.025, 0.25, 0.5, 0.975, 1.0, 1 / 3, 0.9999])
    assert np.array_equal(mcm.pquantile(p, q), orc.quantile(x, q))         # bit-exact given exact order statistics
    mcm._lib.lib().mcm_quantile_use_select(0)                              # full-sort path gives the same bits
    try:
        assert np.array_equal(mcm.pquantile(p, q), orc.quantile(x, q))
    finally:
        mcm._lib.lib().mcm_quantile_use_select(1)
    assert mcm.pquantile(p, 0.5) == orc.quantile(x, 0.5)[0]
    even = mcm.Particles(x[:-1])
    s = orc.sort(x[:-1])                                                   # middle(x, y) = x/2 + y/2 in the particle eltype
    assert mcm.pmedian(even) == float(s[(N - 1) // 2 - 1] / dt(2) + s[(N - 1) // 2] / dt(2)) and mcm.pmedian(p) == orc.sort(x)[N // 2]
    assert np.array_equal(mcm.particles.order_statistics(p, [0, 5, N // 3, N - 1]), orc.sort(x)[[0, 5, N // 3, N - 1]].astype(np.float64))
    with pytest.raises(mcm.ArgumentError):
        mcm.pquantile(p, 1.5)
    xn = x.copy(); xn[5] = np.nan
    with pytest.raises(mcm.ArgumentError):                             # Statistics.quantile: NaN -> ArgumentError
        mcm.pquantile(mcm.Particles(xn), 0.5)
    assert np.isnan(mcm.psort(mcm.Particles(xn)).to_numpy()[-1])       # NaN sorts last
    mcm.particles.check_domain  # flags are sticky; clear for later tests
    bits = C.c_uint32(); mcm._lib.lib().mcm_check_flags(C.byref(bits), 1)


@pytest.mark.parametrize("dt", DTYPES)
def test_systematic_sample_and_constructors(mcm, orc, golden, dt):
    k = golden["known"]
    a = mcm.pm(dt(np.pi), dt(0.1), seed=0)                              # π ± 0.1
    assert a.N == 2000 and a.np_dtype == dt
    tol = 1e-12 if dt == np.float64 else 1e-5
    assert abs(mcm.pstd(a) - k["pstd_pi_pm_0.1"]["value"]) <= tol * 0.1  # README.md:31-32
    # unpermuted grid equals the oracle's within 2 ulp of the underlying z (device Φ⁻¹ < 1 ulp, oracle < 1 ulp)
    s = mcm.systematic_sample(2000, mcm.Normal(dt(np.pi), dt(0.1)), permute=False)
    ref = orc.systematic_normal(2000, np.pi, 0.1, dtype=dt)
    assert np.all(np.diff(s) > 0)                                        # permute=false -> sorted (sampling.jl:4)
    # device z is within 2 ulp(Float64) of the oracle's z (|z| < 4 on this grid); the final μ + σz rounds once more in T
    err = np.abs(s.astype(np.float64) - ref.astype(np.float64))
    assert err.max() <= 0.1 * 2 * 4 * np.finfo(np.float64).eps + np.spacing(dt(np.pi + 0.5))
    # an explicit permutation (e.g. a Julia randperm) is honoured exactly
    perm = orc.randperm(2000, seed=5)
    sp = mcm.systematic_sample(2000, mcm.Normal(dt(np.pi), dt(0.1)), perm=perm)
    assert bits_equal(sp, s[perm])
    # the device permutation is the oracle's Philox permutation bit for bit
    sd = mcm.systematic_sample(2000, mcm.Normal(dt(np.pi), dt(0.1)), permute=True, seed=5)
    assert bits_equal(sd, s[perm])
    # ∓ (N = 100), unary ±, .. (Uniform), Particles(N), Particles()
    b = mcm.mp(dt(2), dt(0.1))
    assert b.N == 100 and round(mcm.pstd(b), 4) == k["two_mp_0.1_pstd_printed"]["value"]
    p100 = mcm.Particles(100)
    assert abs(mcm.pstd(p100) - k["static100_pstd"]["value"]) <= 1e-12 and abs(mcm.pmean(p100)) < 1e-15
    assert mcm.Particles().N == 2000
    assert mcm.prob(mcm.Particles(), "<", 1.0) == k["prob_lt_1"]["value"]   # docs index.md:346-354
    u = mcm.dotdot(1.0, 3.0)
    assert abs(mcm.pmean(u) - 2.0) < 1e-12 and mcm.pminimum(u) > 1.0 and mcm.pmaximum(u) < 3.0
    f = mcm.affine(p100, 2.0, 10.0)                                         # docs index.md:73-74
    assert abs(mcm.pmean(f) - k["normal_fit_2x_plus_10"]["mu"]) <= 1e-13
    assert abs(mcm.pstd(f) - k["normal_fit_2x_plus_10"]["sigma"]) <= 1e-12


def test_norminvcdf_device_vs_truth(mcm, golden):
    p = golden["norminv_p"]
    got = mcm.particles._empty(len(p), mcm._lib.F64)
    pp = mcm.Particles(p)
    mcm._lib.check(mcm._lib.lib().mcm_norminvcdf(mcm._lib.F64, C.c_void_p(pp.ptr), C.c_void_p(got.ptr), len(p)))
    u = ulp_err(got.to_numpy(), golden["norminv_y"])
    assert u.max() <= 1, u.max()     # vs the correctly ROUNDED truth: an error < 1 ulp shows as 0 or 1
    # high-resolution truth (double-double, tests/golden/norminv_truth.npz): grids, both tails down to the
    # subnormals, p next to 1/2 and 1, every branch / piece boundary.  north_star tolerance: 2 ulp.
    import os
    t = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "norminv_truth.npz"))
    got = mcm.particles._empty(len(t["p"]), mcm._lib.F64)
    pp = mcm.Particles(t["p"])
    mcm._lib.check(mcm._lib.lib().mcm_norminvcdf(mcm._lib.F64, C.c_void_p(pp.ptr), C.c_void_p(got.ptr), len(t["p"])))
    u = np.abs((got.to_numpy() - t["y"]) - t["y_lo"]) / np.spacing(np.abs(t["y"]))
    assert u.max() <= 0.75, u.max()  # measured 0.64


@pytest.mark.parametrize("dt", DTYPES)
def test_bootstrap_and_resample_indices_bit_exact(mcm, orc, dt):
    x, y = rand_inputs(dt, 14)
    px, py = mcm.Particles(x), mcm.Particles(y)
    idx = orc.rand_indices(N, N, seed=453)
    assert np.array_equal(mcm.rand_indices(N, N, seed=453).to_numpy(), idx)      # same Philox stream
    assert bits_equal(mcm.bootstrap(px, seed=453).to_numpy(), x[idx])
    b2 = mcm.bootstrap(px, seed=453)
    assert bits_equal(b2.to_numpy(), mcm.bootstrap(px, seed=453).to_numpy())      # runtests.jl:558-562 reproducibility
    n2 = 1234                                                                     # n != N
    assert bits_equal(mcm.bootstrap(px, n2, seed=1).to_numpy(), x[orc.rand_indices(n2, N, seed=1)])
    # array form: ONE index vector shared by all elements (resampling.jl:80-84)
    outs = mcm.bootstrap([px, py, px * 2.0], seed=7)
    i7 = orc.rand_indices(N, N, seed=7)
    assert bits_equal(outs[0].to_numpy(), x[i7]) and bits_equal(outs[1].to_numpy(), y[i7]) and bits_equal(outs[2].to_numpy(), (x * dt(2))[i7])
    # host-supplied indices (e.g. from Julia's rng) are applied verbatim
    mine = np.random.default_rng(0).integers(0, N, size=500)
    assert bits_equal(mcm.bootstrap(px, 500, inds=mine).to_numpy(), x[mine])
    with pytest.raises(IndexError):
        mcm.bootstrap(px, 3, inds=[0, N, 1]); mcm.check_domain()
    # weighted systematic resampling: exact integer scan -> identical indices
    w = np.abs(y.astype(np.float64)) ** 3 + 1e-9
    for u in (0.0, 0.37, 0.999999):
        assert np.array_equal(mcm.resample_indices(w, u).to_numpy(), orc.resample_systematic(w, u)), u
    r = mcm.resample(px, w, 0.37)
    assert bits_equal(r.to_numpy(), x[orc.resample_systematic(w, 0.37)])
    assert np.array_equal(mcm.resample_indices(np.ones(64), 0.5).to_numpy(), np.arange(64))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("M,n", [(3, 1000), (130, 2500), (257, 777), (5, 300001)])
def test_pcov_pcor(mcm, orc, dt, M, n):
    rng = np.random.default_rng(M * 1000 + n)
    X = rng.normal(1.0, 2.0, size=(M, n)).astype(dt)
    X[2] = 2 * X[0] - X[1]                                   # exact linear dependence -> exact off-diagonals
    v = [mcm.Particles(X[j]) for j in range(M)]
    Cg = mcm.pcov(v)
    Cr = orc.cov(X)
    tol = 1e-12 if dt == np.float64 else 1e-5
    assert np.max(np.abs(Cg - Cr)) <= tol * np.max(np.abs(Cr))
    assert np.array_equal(Cg, Cg.T)
    Cu = mcm.pcov(v, corrected=False)
    assert np.max(np.abs(Cu - Cr * (n - 1) / n)) <= tol * np.max(np.abs(Cr))
    R = mcm.pcor(v)
    assert np.allclose(np.diag(R), 1.0) and np.max(np.abs(R - np.corrcoef(X.astype(np.float64)))) <= 10 * tol
    assert mcm.to_matrix(v[:3]).shape == (n, 3)              # Matrix(v) is N x M (runtests.jl:297-322)
    assert abs(mcm.pcov(v[0]) - Cr[0, 0]) <= tol * Cr[0, 0]  # scalar pcov == variance


@pytest.mark.parametrize("dt", DTYPES)
def test_pcov_tcgen05_engine_vs_oracle_and_dmma(mcm, orc, dt):
    """Large problems run on the tcgen05/TMEM int8 path (Ozaki split, exact int32 accumulation); it must agree
    with the oracle and with the FP64 DMMA kernel, and must decline heavy-tailed rows."""
    lib = mcm._lib.lib()
    rng = np.random.default_rng(41)
    for M, n in ((256, 4096), (300, 70001), (640, 20000)):       # second case: 5 flush chunks + row/k padding
        X = rng.normal(3.0, 2.0, size=(M, n)).astype(dt)
        X[2] = 2 * X[0] - X[1]
        v = [mcm.Particles(X[j]) for j in range(M)]
        Ct = mcm.pcov(v)
        assert lib.mcm_gram_last_engine() == 1
        lib.mcm_gram_use_tensor(0)
        try:
            Cd = mcm.pcov(v)
            assert lib.mcm_gram_last_engine() == 0
        finally:
            lib.mcm_gram_use_tensor(1)
        d = np.sqrt(np.diag(Cd))
        tol = 1e-12 if dt == np.float64 else 1e-5
        assert np.max(np.abs(Ct - Cd) / np.outer(d, d)) <= tol
        assert np.array_equal(Ct, Ct.T)
        sub = [0, 1, 2, 7, M // 2, M - 1]
        Cr = orc.cov(X[sub])
        assert np.max(np.abs(Ct[np.ix_(sub, sub)] - Cr)) <= tol * np.max(np.abs(Cr))
    if dt == np.float64:
        # heavy tails stay on the tensor cores: the digit count follows the heaviest row (7 digits up to max/sigma = 160,
        # 8 up to 2560), 1e-12 against the oracle and the DMMA kernel
        cases = {"t(3)": (rng.standard_t(3, size=(256, 20000)), (7, 8)), "lognormal": (np.exp(rng.normal(0.0, 1.5, size=(256, 40001))), (7, 8)),
                 "exp(p) with one wild row": (np.exp(rng.normal(0.0, 0.3, size=(300, 30000))), (7, 8)),
                 "cauchy": (rng.standard_cauchy(size=(256, 16384)), (7, 8))}
        cases["exp(p) with one wild row"][0][17] = rng.standard_t(2, size=30000) * 1e3
        for name, (Xh, digits) in cases.items():
            vh = [mcm.Particles(r) for r in Xh]
            Ct = mcm.pcov(vh)
            if lib.mcm_gram_last_engine() != 1:               # a Cauchy row beyond max/sigma = 2560 may still decline
                assert name == "cauchy"
                continue
            assert lib.mcm_gram_last_digits() in digits, (name, lib.mcm_gram_last_digits())
            lib.mcm_gram_use_tensor(0)
            try:
                Cd = mcm.pcov(vh)
            finally:
                lib.mcm_gram_use_tensor(1)
            d = np.sqrt(np.diag(Cd))
            assert np.max(np.abs(Ct - Cd) / np.outer(d, d)) <= 1e-12, (name, np.max(np.abs(Ct - Cd) / np.outer(d, d)))
            sub = [0, 1, 17, 100, 255]
            Cr = orc.cov(np.ascontiguousarray(Xh[sub]))
            dr = np.sqrt(np.diag(Cr))
            assert np.max(np.abs(Ct[np.ix_(sub, sub)] - Cr) / np.outer(dr, dr)) <= 1e-12, name
        Xi = rng.normal(size=(256, 8192)); Xi[3, 5] = np.inf        # a non-finite row: the guard declines, DMMA propagates it
        mcm.pcov([mcm.Particles(r) for r in Xi])
        assert lib.mcm_gram_last_engine() == 0
    mcm.pcov([mcm.Particles(r) for r in rng.normal(size=(256, 8192)).astype(dt)])
    assert lib.mcm_gram_last_engine() == 1 and lib.mcm_gram_last_digits() == 6
    Xc = rng.normal(size=(256, 8192)); Xc[5] = 1.25                # a constant (zero-variance) row
    Cc = mcm.pcov([mcm.Particles(r) for r in Xc])
    assert np.all(Cc[5] == 0) and np.all(Cc[:, 5] == 0) and abs(Cc[0, 0] - Xc[0].var(ddof=1)) <= 1e-12


def test_pcov_large_offset_is_centred_first(mcm, orc):
    rng = np.random.default_rng(15)
    X = 1e8 + rng.normal(size=(4, 5000))
    Cg = mcm.pcov([mcm.Particles(r) for r in X])
    assert np.max(np.abs(Cg - orc.cov(X))) <= 1e-9            # a raw Gram minus N mu mu' would lose everything


def test_transform_moments_and_mvnormal(mcm):
    # sigmapoints.jl:48-59 doctest: cov(particles) ≈ Σ, mean ≈ m
    m_ = np.array([1.0, -2.0, 0.5]); S = np.array([[3.0, 1, 1], [1, 5, 2], [1, 2, 4]])
    X = [mcm.Particles(500, mcm.Normal(), seed=j + 1) for j in range(3)]
    Y = mcm.transform_moments(X, m_, S)
    assert np.allclose(mcm.pcov(Y), S, atol=1e-10) and np.allclose(mcm.pmean(Y), m_, atol=1e-12)
    P2 = mcm.pm([1.0, 2.0], 0.5)                              # vector μ ± scalar σ -> MvNormal (particles.jl:21)
    assert len(P2) == 2 and P2[0].N == 2000
    assert np.allclose(mcm.pcov(P2), 0.25 * np.eye(2), atol=1e-10)


def test_bymap_prob_and_comparisons(mcm, orc):
    p = mcm.Particles(np.random.default_rng(16).normal(size=200)) + 5.0
    q = mcm.bymap(lambda x: x * x + 1.0, p)
    assert bits_equal(q.to_numpy(), (p * p + 1.0).to_numpy())
    v = mcm.bymap(lambda x, s: np.array([x, s * x]), p, 2.0)          # vector output -> vector of Particles
    assert len(v) == 2 and bits_equal(v[1].to_numpy(), (2.0 * p).to_numpy())
    A = mcm.bymap(lambda x: np.outer([1.0, x], [x, 2.0]), p)          # matrix output
    assert A.shape == (2, 2) and bits_equal(A[1, 0].to_numpy(), (p * p).to_numpy())
    with pytest.raises(mcm.ArgumentError):
        mcm.bymap(lambda x, y: x + y, p, mcm.Particles(np.zeros(5)))   # bymap.jl:34-35
    with pytest.raises(mcm.ArgumentError):
        mcm.bymap(lambda x: x, 1.0)
    xs = p.to_numpy()
    assert mcm.prob(p, "<", 5.2) == np.mean(xs < 5.2) and mcm.prob(p, ">=", 4.0) == np.mean(xs >= 4.0)   # runtests.jl:854-873
    # comparison modes (particles.jl:496-546)
    assert (p > 0) is True and (p < 100) is True
    with pytest.raises(mcm.ComparisonError):
        p < 5.0
    mcm.unsafe_comparisons(True)
    try:
        assert (p < 5.0) == (mcm.pmean(p) < 5.0)
    finally:
        mcm.unsafe_comparisons(False)
    mcm.unsafe_comparisons("montecarlo")
    try:
        assert (p < p + 1.0) is True
        with pytest.raises(mcm.ComparisonError):
            p < 5.0
    finally:
        mcm.unsafe_comparisons("safe")
    assert mcm.approx(p, 5.0) and not mcm.approx(p, 50.0) and mcm.lesssim(p, 50.0) and mcm.gtrsim(p, -50.0)
    assert p == p and not (p == p + 1.0)
    with pytest.raises(mcm.ArgumentError):
        float(p)                                                       # particles.jl:334-338
    assert float(mcm.fill(2.5, 100)) == 2.5


def test_domain_errors_raise_sticky_flag(mcm):
    mcm.check_domain()
    mcm.sqrt(mcm.Particles(np.array([4.0, -1.0, 9.0])))
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()
    mcm.check_domain()                                                # cleared
    mcm.log(mcm.Particles(np.array([1.0, np.nan])))                   # NaN in -> NaN out is NOT a domain error
    mcm.check_domain()
    mcm.fused(lambda a: mcm.log(a) + 1.0)(mcm.Particles(np.array([-2.0, 1.0])))
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()


def test_device_math_is_bit_identical_to_host_twin(mcm, golden):
    """The hand-written Float64 kernels (mcm_fastmath.cuh, mcm_norminv.cuh) compiled for sm_100a with
    -fmad=false must reproduce the host twin bit for bit — which carries the twin's measured ulp bounds
    (tests/test_fastmath_twin.py, millions of points) over to the GPU."""
    from test_fastmath_twin import build_twin, call1
    twin = build_twin()
    rng = np.random.default_rng(21)
    n = 200_000
    k = rng.integers(1, 130000, n).astype(np.float64)
    near = k * (np.pi / 2); near = near + np.spacing(near) * rng.integers(-3, 4, n)
    # fast-path range only: beyond 2^17*pi/2 both sides call their own math library (CUDA vs glibc)
    xs = np.concatenate([rng.uniform(-8, 8, n), np.pi + rng.normal(0, 0.1, n), rng.uniform(-2e5, 2e5, n), near,
                         [0.0, -0.0, np.inf, -np.inf, np.nan, 205886.9, -205886.9]])
    for dev, name in ((mcm.sin, "twin_sin"), (mcm.cos, "twin_cos")):
        assert bits_equal(dev(mcm.Particles(xs)).to_numpy(), call1(twin, name, xs)), name
    xe = np.concatenate([rng.uniform(-708, 708, n), rng.uniform(-5, 5, n), [0.0, 708.0, -708.0, 710.0, -760.0, np.inf, -np.inf, np.nan]])
    assert bits_equal(mcm.exp(mcm.Particles(xe)).to_numpy(), call1(twin, "twin_exp", xe))
    xp = np.concatenate([np.abs(np.pi + rng.normal(0, 0.3, n)), np.exp(rng.uniform(-60, 60, n))])
    for y in (1.5, 0.5, 2.5, -0.5, -1.5, 7.5):
        e = np.full(xp.size, y); out = np.empty_like(xp)
        twin.twin_pow(xp.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(xp.size))
        assert bits_equal((mcm.Particles(xp) ** y).to_numpy(), out), y
    p = golden["norminv_p"]
    got = mcm.particles._empty(len(p), mcm._lib.F64)
    mcm._lib.check(mcm._lib.lib().mcm_norminvcdf(mcm._lib.F64, C.c_void_p(mcm.Particles(p).ptr), C.c_void_p(got.ptr), len(p)))
    assert bits_equal(got.to_numpy(), call1(twin, "twin_norminv", p))
    mcm._lib.lib().mcm_check_flags(C.byref(C.c_uint32()), 1)
