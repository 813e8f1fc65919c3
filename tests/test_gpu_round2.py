"""Round-2 additions on the device path, each against the oracle / the reference's own assertions:
RoundingMode forms of round / rem / div, rem2pi / mod2pi, round(Int, p), the all-reductions
(!, isinteger, iszero, isfinite) — src/particles.jl:345-359,581-585 pinned by test/runtests.jl:466-483 —
plus the advisor's findings (parallel loop write-back, floored //, gather bounds, == on infinities)."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal, ulp_err

pytestmark = pytest.mark.gpu
DTYPES = [np.float64, np.float32]
N = 20000


def inputs(dt, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(0.3, 2.0, N).astype(dt)
    x[:12] = [0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.49999997, -0.49999997, 3.0, -7.0]
    y = rng.normal(-0.2, 1.5, N).astype(dt)
    y[np.abs(y) < 0.05] = dt(0.7)
    return x, y


@pytest.mark.parametrize("dt", DTYPES)
def test_round_with_rounding_modes_bit_exact(mcm, orc, dt):
    x, _ = inputs(dt)
    p = mcm.Particles(x)
    for mode, name in ((mcm.RoundNearest, "ROUND"), (mcm.RoundToZero, "TRUNC"), (mcm.RoundUp, "CEIL"),
                       (mcm.RoundDown, "FLOOR"), (mcm.RoundNearestTiesAway, "ROUND_TIESAWAY"),
                       (mcm.RoundNearestTiesUp, "ROUND_TIESUP"), (mcm.RoundFromZero, "ROUND_FROMZERO")):
        assert bits_equal(mcm.round_(p, mode).to_numpy(), orc.unary(name, x)), mode
    # the definitions themselves (Base): ties away from zero / towards +Inf / away from zero
    t = mcm.Particles(np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.2, -0.2], dtype=dt))
    assert list(mcm.round_(t, mcm.RoundNearestTiesAway).to_numpy()) == [1, 2, 3, -1, -2, -3, 0, -0.0]
    assert list(mcm.round_(t, mcm.RoundNearestTiesUp).to_numpy()) == [1, 2, 3, -0.0, -1, -2, 0, -0.0]
    assert list(mcm.round_(t, mcm.RoundFromZero).to_numpy()) == [1, 2, 3, -1, -2, -3, 1, -1]


@pytest.mark.parametrize("dt", DTYPES)
def test_rem_div_with_rounding_modes_bit_exact(mcm, orc, dt):
    x, y = inputs(dt, 1)
    px, py = mcm.Particles(x), mcm.Particles(y)
    for mode, rn, dn in ((mcm.RoundToZero, "REM", "FDIV"), (mcm.RoundNearest, "REM_NEAREST", "DIV_NEAREST"),
                         (mcm.RoundDown, "REM_DOWN", "DIV_DOWN"), (mcm.RoundUp, "REM_UP", "DIV_UP"),
                         (mcm.RoundFromZero, "REM_FROMZERO", "DIV_FROMZERO")):
        assert bits_equal(mcm.rem(px, py, mode).to_numpy(), orc.binary(rn, x, y)), mode
        assert bits_equal(mcm.div(px, py, mode).to_numpy(), orc.binary(dn, x, y)), mode
        assert bits_equal(mcm.rem(px, 1.5, mode).to_numpy(), orc.binary(rn, x, dt(1.5))), mode   # runtests.jl:477,479
        assert bits_equal(mcm.div(px, 1.5, mode).to_numpy(), orc.binary(dn, x, dt(1.5))), mode
    # runtests.jl:474-475,478: rem(p, p) == 0, div(p, p) == 1 for every mode that divides exactly
    nz = mcm.Particles(np.abs(x) + dt(0.1))
    for mode in (mcm.RoundToZero, mcm.RoundFromZero, mcm.RoundDown):
        assert mcm.iszero(mcm.rem(nz, nz, mode)) and mcm.div(nz, nz, mode) == mcm.fill(1.0, N, dt)
    # the identity x == y*div + rem holds for every mode up to the rounding of the product
    for mode in (mcm.RoundToZero, mcm.RoundNearest, mcm.RoundDown, mcm.RoundUp, mcm.RoundFromZero):
        q, r = mcm.div(px, py, mode).to_numpy().astype(np.float64), mcm.rem(px, py, mode).to_numpy().astype(np.float64)
        assert np.allclose(q * y + r, x, rtol=4 * np.finfo(dt).eps, atol=4 * np.finfo(dt).eps)
        assert np.all(np.abs(r) <= np.abs(y) * (1 + 1e-6))


@pytest.mark.parametrize("dt", DTYPES)
def test_python_floordiv_is_floored_and_pairs_with_mod(mcm, dt):
    x, y = inputs(dt, 2)
    px, py = mcm.Particles(x), mcm.Particles(y)
    q, r = (px // py).to_numpy().astype(np.float64), (px % py).to_numpy().astype(np.float64)
    assert np.allclose(q * y + r, x, rtol=4 * np.finfo(dt).eps, atol=4 * np.finfo(dt).eps)
    a = mcm.Particles(np.array([-7.0, 7.0, -7.0, 7.5], dtype=dt))
    assert list((a // 2.0).to_numpy()) == [-4, 3, -4, 3] and list((a % 2.0).to_numpy()) == [1, 1, 1, 1.5]
    assert list((7.0 // mcm.Particles(np.array([2.0, -2.0], dtype=dt))).to_numpy()) == [3, -4]     # __rfloordiv__
    assert list(mcm.div(a, 2.0).to_numpy()) == [-3, 3, -3, 3]                                      # Julia's ÷ truncates


def test_rem2pi_and_mod2pi_vs_exact_truth(mcm, orc):
    rng = np.random.default_rng(3)
    k = rng.integers(1, 10 ** 6, 300).astype(float)
    near = k * (np.pi / 2); near = near + np.spacing(near) * rng.integers(-2, 3, 300)
    x = np.concatenate([rng.normal(0, 3, 600), rng.uniform(-2e5, 2e5, 400), 10.0 ** rng.uniform(5, 300, 300) * rng.choice([-1, 1], 300),
                        near, -near, [0.0, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 1e22, -1e22, 1.7e308, 205887.0, 205888.0]])
    p = mcm.Particles(x)
    for mode in (mcm.RoundNearest, mcm.RoundToZero, mcm.RoundDown, mcm.RoundUp):
        got = mcm.rem2pi(p, mode).to_numpy()
        ref = orc.rem2pi_truth(x, mode)
        assert ulp_err(got, ref).max() <= 1, mode          # correctly rounded on this set (measured 0 differences on the twin)
        assert (got != ref).mean() < 1e-3, mode
    assert bits_equal(mcm.mod2pi(p).to_numpy(), mcm.rem2pi(p, mcm.RoundDown).to_numpy())           # runtests.jl:476
    m = mcm.mod2pi(p).to_numpy()
    assert np.all((m >= 0) & (m <= 2 * np.pi))
    # Float32 goes through Float64 like Base
    xf = x[:1000].astype(np.float32)
    gf = mcm.mod2pi(mcm.Particles(xf)).to_numpy()
    assert np.array_equal(gf, orc.rem2pi_truth(xf.astype(np.float64), "RoundDown").astype(np.float32))
    # Inf is a DomainError in Base
    mcm.rem2pi(mcm.Particles(np.array([np.inf, 1.0])), mcm.RoundNearest)
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()


@pytest.mark.parametrize("dt", DTYPES)
def test_round_int_and_all_reductions(mcm, dt):
    x, _ = inputs(dt, 4)
    p = mcm.Particles(x)
    for mode, f in ((mcm.RoundNearest, np.rint), (mcm.RoundToZero, np.trunc), (mcm.RoundUp, np.ceil), (mcm.RoundDown, np.floor),
                    (mcm.RoundFromZero, lambda v: np.where(np.signbit(v), np.floor(v), np.ceil(v)))):
        q = mcm.round_int(p, mode)
        assert isinstance(q, mcm.IntParticles) and q.N == N and q.to_numpy().dtype == np.int64       # runtests.jl:472
        assert np.array_equal(q.to_numpy(), f(x.astype(np.float64)).astype(np.int64)), mode          # runtests.jl:473
    assert mcm.round_int(mcm.Particles((0.001 * x / 10).astype(dt))) == 0                            # runtests.jl:471
    assert not (mcm.round_int(p) == 0)
    # InexactError cases raise the domain flag
    mcm.check_domain()
    mcm.round_int(mcm.Particles(np.array([1.0, np.nan, 1e30], dtype=dt)))
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()
    # runtests.jl:464-469: isfinite(p), iszero(0p), iszero(p, 0.1), !iszero(p), !(!p), !(0p)
    small = mcm.Particles((x * dt(1e-3)).astype(dt))
    zero = p * 0
    assert mcm.isfinite(p) and mcm.iszero(zero) and mcm.iszero(small, 0.1) and not mcm.iszero(p)
    assert not mcm.not_(p) and mcm.not_(zero)
    assert mcm.isinteger(mcm.round_(p)) and not mcm.isinteger(p) and mcm.isinteger(zero)
    assert not mcm.isinteger(mcm.Particles(np.array([1.0, np.inf], dtype=dt)))
    assert not mcm.isfinite(mcm.Particles(np.array([1.0, np.inf], dtype=dt)))


def test_equality_compares_particle_vectors(mcm):
    a = mcm.Particles(np.array([1.0, np.inf, -np.inf, 0.0]))
    b = mcm.Particles(np.array([1.0, np.inf, -np.inf, -0.0]))
    assert a == b                                                # Inf == Inf, 0.0 == -0.0 (x - y == 0 failed on Inf)
    assert not (a == mcm.Particles(np.array([1.0, np.inf, np.inf, 0.0])))
    n = mcm.Particles(np.array([1.0, np.nan]))
    assert not (n == n)                                          # NaN is never equal (Base == on vectors)
    mcm.unsafe_comparisons("montecarlo")
    try:
        x = mcm.Particles(np.array([1.0, 2.0, 3.0])); y = mcm.Particles(np.array([1.5, 2.5, 3.5]))
        assert x < y and not (y < x) and x <= y and 0.5 < x and x < 3.5
        with pytest.raises(mcm.ComparisonError):
            x < mcm.Particles(np.array([0.5, 2.5, 3.5]))
    finally:
        mcm.unsafe_comparisons("safe")


def test_out_of_range_indices_raise_and_write_nan(mcm):
    p = mcm.Particles(np.arange(10.0))
    q = mcm.Particles(np.arange(10.0) * 2)
    with pytest.raises(IndexError):
        mcm.bootstrap(p, 4, inds=[0, 3, 10, 2])
    with pytest.raises(IndexError):
        mcm.bootstrap([p, q], 4, inds=[0, -1, 9, 2])
    r = mcm.bootstrap([p, q], 4, inds=[0, 3, 9, 2])              # the flag was cleared: valid calls work again
    assert list(r[0].to_numpy()) == [0, 3, 9, 2] and list(r[1].to_numpy()) == [0, 6, 18, 4]
    # the raw C-ABI call leaves NaN (not garbage) in the skipped outputs
    L = mcm._lib
    idx = mcm.particles._upload_i64(np.array([1, 99, 2], dtype=np.int64))
    out = mcm.particles._empty(3, L.F64)
    L.check(L.lib().mcm_gather(L.F64, C.c_void_p(p.ptr), C.c_void_p(idx.ptr), C.c_void_p(out.ptr), 3, 10))
    got = out.to_numpy()
    assert got[0] == 1 and np.isnan(got[1]) and got[2] == 2
    L.check(L.lib().mcm_clear_flags(0x80000000))
    mcm.check_domain()                                           # nothing else is pending


@pytest.mark.parametrize("dt", DTYPES)
def test_loop_tape_parallel_writeback_on_device(mcm, dt):
    """(a, b) -> (a + b, a) iterated in registers: each step must read the OLD a (tape.py)."""
    from mcm_b200 import tape
    L = mcm._lib
    rng = np.random.default_rng(6)
    a0, b0 = rng.normal(size=4096).astype(dt), rng.normal(size=4096).astype(dt)
    for step, ref in ((lambda a, b: (a + b, a), lambda a, b: (a + b, a)), (lambda a, b: (b, a), lambda a, b: (b, a))):
        t = tape.trace_loop(step, 2, 0)
        pa, pb = mcm.Particles(a0), mcm.Particles(b0)
        oa, ob = mcm.particles._empty(4096, pa.dtype), mcm.particles._empty(4096, pa.dtype)
        tape.run_tape(t, [pa.ptr, pb.ptr], 4096, pa.dtype, [oa.ptr, ob.ptr], reps=7)
        wa, wb = a0.copy(), b0.copy()
        for _ in range(7):
            wa, wb = ref(wa, wb)
        assert bits_equal(oa.to_numpy(), wa) and bits_equal(ob.to_numpy(), wb)


# ---- engine 3: kernels generated from the tape at run time (csrc/chain_jit.cu) ----------------------------------
def _run(mcm, t, arrays, n, dt_code, engine, reps=0, stats=True):
    from mcm_b200 import tape
    L = mcm._lib
    L.check(L.lib().mcm_chain_force_engine(engine))
    try:
        ins = [mcm.Particles(a) for a in arrays]
        outs = [mcm.particles._empty(n, dt_code) for _ in range(t.nout)]
        st = tape.run_tape(t, [p.ptr for p in ins], n, dt_code, [o.ptr for o in outs], stats_output=0 if stats else None, reps=reps)
        eng = L.lib().mcm_chain_last_engine()
        return [o.to_numpy() for o in outs], st, eng
    finally:
        L.check(L.lib().mcm_chain_force_engine(0))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [1, 3, 1001, 262144 + 5])
def test_generated_kernels_are_bit_identical_to_the_interpreter(mcm, dt, n):
    from mcm_b200 import chains, tape
    L = mcm._lib
    code = L.F64 if dt == np.float64 else L.F32
    rng = np.random.default_rng(n)
    a = (np.pi + 0.1 * rng.normal(size=n)).astype(dt)
    b = rng.normal(0.5, 1.0, n).astype(dt)
    cases = [
        (tape.trace(chains.chain10, 1), [a], 0),
        (tape.trace(chains.readme_expr, 1), [a], 0),
        # constant-exponent powers of several kinds, integer powers, two outputs, ternary ops, a user-style chain
        (tape.trace(lambda x, y: (mcm.sqrt(abs(x)) ** 2.5 + abs(y) ** 0.5 - (abs(x) + 1.0) ** -1.5 + x ** 5 + y ** -3,
                                  mcm.clamp(mcm.muladd(x, y, 0.25), -2.0, 2.0) / (abs(y) ** 1.7 + 1.0)), 2), [a, b], 0),
        (tape.trace(lambda x, y: mcm.log(abs(x * y) + 0.1) + mcm.cos(x - y) * mcm.atan(y, x) + mcm.mod2pi(x * 7.0), 2), [a, b], 0),
        (tape.trace_loop(lambda u, v: (u + 0.01 * v, v - 0.01 * mcm.sin(u)), 2, 0), [a, b], 37),
    ]
    for t, arrs, reps in cases:
        o3, s3, e3 = _run(mcm, t, arrs, n, code, 3, reps)
        o2, s2, e2 = _run(mcm, t, arrs, n, code, 2, reps)
        assert e3 == 3 and e2 == 2
        for x3, x2 in zip(o3, o2):
            assert bits_equal(x3, x2)
        assert s3[0] == s2[0] == n and s3[3] == s2[3] and s3[4] == s2[4]                    # n, min, max exact
        assert abs(s3[1] - s2[1]) <= 1e-13 * max(abs(s2[1]), 1e-300) + 1e-300
        assert abs(s3[2] - s2[2]) <= 1e-12 * s2[2] + 1e-300
    mcm.check_domain()


@pytest.mark.parametrize("opts", [(2, 1, 2, 0), (2, 0, 4, 0), (4, 0, 1, 0), (3, 0, 2, 1), (2, 0, 4, 1), (4, 0, 1, 1), (0, 0, 0, 2)])
def test_generation_options_do_not_change_results(mcm, opts):
    """CTAs per SM / prefetch / packs per thread / input staging through shared memory (bulk async copies) are tunables
    of the generator: same bits, any size (tiles, leftover packs, ragged last pack), one and two inputs."""
    from mcm_b200 import chains, tape
    L = mcm._lib
    lib = L.lib()
    t = tape.trace(chains.chain10, 1)
    t2 = tape.trace_loop(lambda u, v: (u + 0.01 * v, v - 0.01 * mcm.sin(u)), 2, 0)
    t3 = tape.trace(lambda u, v: (mcm.sqrt(abs(u)) / (v * v + 1.0), u - v), 2)

    def setopts(o):
        for which, val in zip((0, 1, 2, 4), o):
            L.check(lib.mcm_chain_jit_option(which, val))
    try:
        for n in (1, 2, 5, 7, 9, 4099, 100003, 1 << 20):
            rng = np.random.default_rng(n)
            a = np.pi + 0.1 * rng.normal(size=n); b = rng.normal(size=n)
            setopts((3, 0, 1, 0))
            o_ref, s_ref, _ = _run(mcm, t, [a], n, L.F64, 3)
            l_ref, _, _ = _run(mcm, t2, [a, b], n, L.F64, 3, reps=11)
            m_ref, _, _ = _run(mcm, t3, [a, b], n, L.F64, 3)
            f_ref, _, _ = _run(mcm, t3, [a.astype(np.float32), b.astype(np.float32)], n, L.F32, 3)
            setopts(opts)
            o, s, e = _run(mcm, t, [a], n, L.F64, 3)
            l, _, _ = _run(mcm, t2, [a, b], n, L.F64, 3, reps=11)
            m, _, _ = _run(mcm, t3, [a, b], n, L.F64, 3)
            f, _, _ = _run(mcm, t3, [a.astype(np.float32), b.astype(np.float32)], n, L.F32, 3)
            assert e == 3 and bits_equal(o[0], o_ref[0]) and bits_equal(l[0], l_ref[0]) and bits_equal(l[1], l_ref[1])
            assert bits_equal(m[0], m_ref[0]) and bits_equal(m[1], m_ref[1]) and bits_equal(f[0], f_ref[0]) and bits_equal(f[1], f_ref[1])
            assert s[0] == n and s[3] == s_ref[3] and s[4] == s_ref[4] and abs(s[1] - s_ref[1]) <= 1e-13 * abs(s_ref[1])
            assert abs(s[2] - s_ref[2]) <= 1e-12 * s_ref[2] + 1e-300
    finally:
        setopts((0, 0, 0, 2))


def test_generated_kernel_replays_out_of_range_packs_exactly(mcm):
    """Arguments outside the branch-free fast paths (|x| >= 708 for exp, >= 2^17 pi/2 for sin, bases outside
    (1e-280, 1e280) or negative for x^1.5, NaN, Inf) make the pack replay through the exact instantiation."""
    from mcm_b200 import tape
    L = mcm._lib
    x = np.array([0.5, 800.0, -750.0, 3e5, -1e300, 1e-300, 1e290, 0.0, -0.0, np.inf, -np.inf, np.nan, 2.0, 708.0, 205887.0, 1e-285,
                  709.5, -708.0000001, 1.0, 4.0], dtype=np.float64)
    x = np.concatenate([x, np.random.default_rng(0).normal(3.0, 0.5, 4096 - len(x))])
    for f in (lambda p: mcm.exp(p) + mcm.sin(p), lambda p: p ** 1.5, lambda p: mcm.exp(-p) * mcm.cos(p) + abs(p) ** 2.5,
              lambda p: mcm.sqrt(p) + mcm.log(p)):
        t = tape.trace(f, 1)
        o3, s3, e3 = _run(mcm, t, [x], len(x), L.F64, 3)
        o2, s2, e2 = _run(mcm, t, [x], len(x), L.F64, 2)
        assert e3 == 3 and bits_equal(o3[0], o2[0])
        b3 = C.c_uint32(); L.check(L.lib().mcm_check_flags(C.byref(b3), 1))
    # the DomainError flag survives the fast path: sqrt(-1) inside a generated kernel
    t = tape.trace(lambda p: mcm.sqrt(p - 10.0), 1)
    _run(mcm, t, [x], len(x), L.F64, 3)
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()


def test_generation_policy_and_cache(mcm):
    from mcm_b200 import tape
    L = mcm._lib
    lib = L.lib()
    f = lambda p: mcm.tanh(p) * 1.25 + mcm.sqrt(abs(p)) ** 3.5
    t = tape.trace(f, 1)
    small = np.linspace(-1, 1, 2000)
    before = lib.mcm_chain_jit_compiles()
    _, _, e = _run(mcm, t, [small], len(small), L.F64, 0)
    assert e == 2 and lib.mcm_chain_jit_compiles() == before          # small launch: the interpreter, nothing compiled
    big = np.linspace(-1, 1, 1 << 20)
    o_big, _, e = _run(mcm, t, [big], len(big), L.F64, 0)
    assert e == 3 and lib.mcm_chain_jit_compiles() == before + 1      # large launch: generated once ...
    _, _, e = _run(mcm, t, [big], len(big), L.F64, 0)
    assert e == 3 and lib.mcm_chain_jit_compiles() == before + 1      # ... and cached
    _, _, e = _run(mcm, t, [small], len(small), L.F64, 0)
    assert e == 3                                                      # a cached kernel serves small launches too
    # a different run-time constant reuses the kernel; a different pow exponent is a different kernel
    t2 = tape.trace(lambda p: mcm.tanh(p) * 7.5 + mcm.sqrt(abs(p)) ** 3.5, 1)
    _, _, e = _run(mcm, t2, [big], len(big), L.F64, 0)
    assert e == 3 and lib.mcm_chain_jit_compiles() == before + 1
    t3 = tape.trace(lambda p: mcm.tanh(p) * 1.25 + mcm.sqrt(abs(p)) ** 2.5, 1)
    _, _, e = _run(mcm, t3, [big], len(big), L.F64, 0)
    assert e == 3 and lib.mcm_chain_jit_compiles() == before + 2
    L.check(lib.mcm_chain_jit_mode(0))
    try:
        t4 = tape.trace(lambda p: mcm.tanh(p) * 1.25 - 3.0, 1)
        _, _, e = _run(mcm, t4, [big], len(big), L.F64, 0)
        assert e == 2
    finally:
        L.check(lib.mcm_chain_jit_mode(1))
    ref = np.tanh(big) * 1.25 + np.sqrt(np.abs(big)) ** 3.5
    assert np.allclose(o_big[0], ref, rtol=1e-14, atol=1e-15)


# ---- any univariate Distribution on device (src/sampling.jl:10, src/types.jl:67-90, src/particles.jl:42,50) -------------
def _dist_cases(mcm):
    return [("NORMAL", mcm.Normal(0.5, 2.0), (0.5, 2.0)), ("UNIFORM", mcm.Uniform(-1.0, 3.0), (-1.0, 3.0)),
            ("EXPONENTIAL", mcm.Exponential(2.5), (2.5,)), ("LOGNORMAL", mcm.LogNormal(0.3, 0.7), (0.3, 0.7)),
            ("CAUCHY", mcm.Cauchy(1.0, 2.0), (1.0, 2.0)), ("LAPLACE", mcm.Laplace(1.0, 2.0), (1.0, 2.0)),
            ("LOGISTIC", mcm.Logistic(1.0, 2.0), (1.0, 2.0)), ("RAYLEIGH", mcm.Rayleigh(1.5), (1.5,)),
            ("WEIBULL", mcm.Weibull(1.7, 2.0), (1.7, 2.0)), ("PARETO", mcm.Pareto(3.0, 2.0), (3.0, 2.0)),
            ("GUMBEL", mcm.Gumbel(1.0, 2.0), (1.0, 2.0)), ("TRIANGULAR", mcm.TriangularDist(1.0, 4.0, 2.0), (1.0, 4.0, 2.0)),
            ("FRECHET", mcm.Frechet(2.5, 1.5), (2.5, 1.5)), ("KUMARASWAMY", mcm.Kumaraswamy(2.0, 3.0), (2.0, 3.0)),
            ("GAMMA", mcm.Gamma(0.05, 1.5), (0.05, 1.5)), ("GAMMA", mcm.Gamma(7.3, 1.5), (7.3, 1.5)),
            ("GAMMA", mcm.Gamma(5000.0, 1.0), (5000.0, 1.0)), ("CHISQ", mcm.Chisq(3.0), (3.0,)),
            ("INVERSEGAMMA", mcm.InverseGamma(3.0, 2.0), (3.0, 2.0)), ("BETA", mcm.Beta(0.5, 0.5), (0.5, 0.5)),
            ("BETA", mcm.Beta(30.0, 2.0), (30.0, 2.0)), ("BETA", mcm.Beta(200.0, 300.0), (200.0, 300.0)),
            ("TDIST", mcm.TDist(1.0), (1.0,)), ("TDIST", mcm.TDist(2.5), (2.5,)), ("TDIST", mcm.TDist(200.0), (200.0,)),
            ("FDIST", mcm.FDist(3.0, 7.0), (3.0, 7.0)), ("FDIST", mcm.FDist(10.0, 4.0), (10.0, 4.0))]


def test_device_quantiles_match_the_host_twin_on_the_systematic_grid(mcm):
    """systematic_sample(N, d; permute=false) = quantile.(d, (2i-1)/(2N)) for every device family: the device runs the
    same source as the host twin, which tests/test_dist_twin.py pins to scipy / mpmath at 1e-12; the two differ only by
    libm vs libdevice log/exp/pow/lgamma (tolerance tier, 1e-12 relative to max(|x|, IQR/1000))."""
    from test_dist_twin import build_dist_twin, grid, twin_quantile
    twin = build_dist_twin()
    for N in (2000, 100003):
        p, q = grid(N)
        for kind, d, params in _dist_cases(mcm):
            got = mcm.systematic_sample(N, d, permute=False)
            ref = twin_quantile(twin, kind, params, p, q)
            iqr = abs(ref[(3 * N) // 4] - ref[N // 4])
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * iqr)
            assert got.shape == (N,) and not np.isnan(got).any() and err.max() <= 1e-12, (kind, params, N, err.max(), int(np.argmax(err)))
            assert np.all(np.diff(got) >= 0), (kind, params)
    mcm.check_domain()


def test_systematic_sample_of_other_families_like_the_reference_tests(mcm):
    """test/runtests.jl:24-28: moment fits of systematic samples; a permuted sample is a permutation of the grid;
    Float32 eltype comes from the parameters (types.jl:85-90)."""
    x = mcm.systematic_sample(10000, mcm.Normal(1, 1))
    assert abs(x.mean() - 1) < 0.1 and abs(x.std() - 1) < 0.1
    g = mcm.systematic_sample(10000, mcm.Gamma(1, 1))
    assert abs(g.mean() - 1) < 0.02 and g.min() > 0
    t = mcm.systematic_sample(10000, mcm.TDist(1))
    assert np.isfinite(t).all() and abs(np.median(t)) < 1e-3
    b = mcm.systematic_sample(10000, mcm.Beta(1, 1))                     # Beta(1,1) = Uniform(0,1): fit gives (1,1)
    m, v = b.mean(), b.var()
    k = m * (1 - m) / v - 1
    assert abs(m * k - 1) < 0.1 and abs((1 - m) * k - 1) < 0.1
    s0 = mcm.systematic_sample(5000, mcm.Gamma(2.0, 3.0), permute=False)
    s1 = mcm.systematic_sample(5000, mcm.Gamma(2.0, 3.0), permute=True, seed=3)
    assert not np.array_equal(s0, s1) and np.array_equal(np.sort(s1), s0)
    f = mcm.systematic_sample(1000, mcm.Exponential(np.float32(2.0)), permute=False, as_particles=True)
    assert f.to_numpy().dtype == np.float32
    with pytest.raises(mcm.ArgumentError):
        mcm.Gamma(-1.0, 1.0)


def test_boxplus_boxtimes_outer_product(mcm):
    """runtests.jl:43-44 (`-50 ⊞ Normal(0,1) ≈ -50 ± 1`, `10 ⊠ Normal(0,1) ≈ 10*Particles(Normal(0,1))`) and
    runtests.jl:610-632 (outer_product: d variables, <= N particles, identity covariance, reproducible by seed)."""
    a = mcm.boxplus(-50, mcm.Normal(0, 1), seed=1)
    assert abs(mcm.pmean(a) + 50) < 1e-9 and abs(mcm.pstd(a) - 1) < 2e-3
    assert mcm.approx(a, mcm.pm(-50, 1))
    b = mcm.boxtimes(10, mcm.Normal(0, 1), seed=1)
    assert abs(mcm.pmean(b)) < 1e-9 and abs(mcm.pstd(b) - 10) < 2e-2
    g = mcm.boxtimes(2.0, mcm.Gamma(3.0, 1.0))
    assert abs(mcm.pmean(g) - 6.0) < 1e-2 and abs(mcm.pvar(g) - 12.0) < 0.2
    p = mcm.otimes([0.0, 0.0], [1.0, 1.0])
    assert len(p) == 2 and p[0].N <= 100_000 and p[0].N == 316 ** 2
    assert np.allclose(mcm.pcov(p), np.eye(2), atol=1e-1)
    p1 = mcm.outer_product([mcm.Normal(0, 1), mcm.Gamma(2.0, 1.0), mcm.Uniform(0, 1)], 1000, seed=38)
    p2 = mcm.outer_product([mcm.Normal(0, 1), mcm.Gamma(2.0, 1.0), mcm.Uniform(0, 1)], 1000, seed=38)
    assert len(p1) == 3 and p1[0].N == 1000 and all(x == y for x, y in zip(p1, p2))
    c = mcm.pcov(p1)
    assert abs(c[0, 1]) < 1e-12 and abs(c[0, 2]) < 1e-12 and abs(c[1, 2]) < 1e-12      # a full product grid is uncorrelated


def test_iid_construction_and_quantile_vector_on_device(mcm):
    """Particles(N, d; systematic=false) (types.jl:71,79): inversion of Philox uniforms on device; reproducible by
    seed, right law (Kolmogorov distance), upper tail not truncated.  quantile.(d, p) on a probability vector
    equals the systematic grid when p is the grid."""
    import scipy.stats as st
    n = 200000
    for d, ref in ((mcm.Normal(1.0, 2.0), st.norm(1.0, 2.0)), (mcm.Gamma(2.5, 1.5), st.gamma(2.5, scale=1.5)),
                   (mcm.Exponential(0.5), st.expon(scale=0.5)), (mcm.Beta(2.0, 5.0), st.beta(2.0, 5.0)),
                   (mcm.TDist(4.0), st.t(4.0))):
        p1 = mcm.Particles(n, d, systematic=False, seed=7)
        p2 = mcm.Particles(n, d, systematic=False, seed=7)
        p3 = mcm.Particles(n, d, systematic=False, seed=8)
        x = p1.to_numpy()
        assert np.array_equal(x, p2.to_numpy()) and not np.array_equal(x, p3.to_numpy())
        assert st.kstest(x, ref.cdf).statistic < 1.63 / np.sqrt(n) * 1.5, type(d).__name__
    N = 4001
    pg = (2 * np.arange(N) + 1) / (2.0 * N)
    for d in (mcm.Gamma(2.5, 1.5), mcm.LogNormal(0.0, 0.5), mcm.Normal(0.0, 1.0)):
        qv = mcm.quantile_dist(d, pg).to_numpy()
        sv = mcm.systematic_sample(N, d, permute=False)
        lo = pg < 0.9                                    # 1 - p is rounded once in the vector form: compare off the far tail
        assert np.allclose(qv[lo], sv[lo], rtol=1e-12, atol=0) and np.allclose(qv, sv, rtol=1e-9)
    edge = mcm.quantile_dist(mcm.Exponential(1.0), np.array([0.0, 1.0, -0.5, 2.0])).to_numpy()
    assert edge[0] == 0.0 and edge[1] == np.inf and np.isnan(edge[2:]).all()
    with pytest.raises(mcm.DomainError):                 # p outside [0,1]: NaN + the sticky flag (Distributions throws)
        mcm.check_domain()


def test_host_supplied_quantile_table_and_discrete_distributions(mcm):
    """Any other Distribution goes through a host-supplied quantile table (GenericDistribution; a scipy frozen
    distribution works as is).  Discrete ones are sampled iid by default (systematic = VS == Continuous, types.jl:67;
    runtests.jl:274-277: Particles(Poisson(50)))."""
    import scipy.stats as st
    d = st.skewnorm(4.0, loc=1.0, scale=2.0)
    s = mcm.systematic_sample(3000, d, permute=False)
    N = 3000
    assert np.array_equal(s, d.ppf((2.0 * np.arange(N) + 1.0) / (2.0 * N)))
    sp = mcm.systematic_sample(3000, d, permute=True, seed=5)
    assert np.array_equal(np.sort(sp), s) and not np.array_equal(sp, s)
    p = mcm.Particles(2000, st.poisson(50))
    x = p.to_numpy()
    assert np.all(x == np.round(x)) and abs(x.mean() - 50) < 1.0 and abs(x.var() - 50) < 8.0
    y = (p * p - 1.0).to_numpy()
    assert np.all(y == np.round(y))
    c = mcm.boxplus(1, st.binom(3, 0.5))
    assert set(np.unique(c.to_numpy())) <= {1.0, 2.0, 3.0, 4.0}


def _wide_doubles(n, seed):
    """Doubles with random sign, exponents over the whole range (denormals, zeros, Inf, NaN included) and random
    mantissas, with a dense share around 1 and values hugging the window edges of the inline sqrt / division."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 1 << 63, n, dtype=np.int64) | (rng.integers(0, 2, n, dtype=np.int64) << 63)
    x = bits.view(np.float64).copy()
    k = n // 4
    x[:k] = rng.normal(0.0, 3.0, k)                                   # ordinary data
    x[k:2 * k] = np.exp2(rng.uniform(-60, 60, k)) * rng.choice([-1.0, 1.0], k)
    edge = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1.0, -1.0,
                     2.0 ** -400, 2.0 ** 400, 2.0 ** -401, 2.0 ** 401, 2.0 ** -900, 2.0 ** 900, 2.0 ** 901, 2.0 ** -901,
                     np.nextafter(2.0 ** 400, np.inf), np.nextafter(2.0 ** -400, 0.0), 2.0 ** 498, 2.0 ** 499, 2.0 ** -498, 2.0 ** -499,
                     1e205, 1e250, 1e-250, 4.0, 9.0, 1e10, 1e16, 8.2e5, 8.4e5, 3e9, -1e10, 205886.9, 205887.1, 1e22])
    x[2 * k:2 * k + len(edge)] = edge
    return x


def test_inline_sqrt_and_division_are_the_ieee_results(mcm):
    """Generated kernels expand sqrt, /, inv and x^(k/2) inline (compiler's fast-path sequences restated, no slow-path
    call); values outside the windows replay through the real operators.  Every result must be the correctly rounded
    one: bit-equal to numpy's sqrt and / on the host, for operands over the whole double range."""
    from mcm_b200 import tape
    L = mcm._lib
    n = 1 << 20
    x, y = _wide_doubles(n, 1), _wide_doubles(n, 2)
    with np.errstate(all="ignore"):
        want = {"sqrt": np.sqrt(x), "div": x / y, "inv": 1.0 / x, "cdiv": 3.0 / y, "divc": x / 7.0}
    got = {}
    got["sqrt"], _, e1 = _run(mcm, tape.trace(lambda p: mcm.sqrt(p), 1), [x], n, L.F64, 3, stats=False)
    got["div"], _, e2 = _run(mcm, tape.trace(lambda p, q: p / q, 2), [x, y], n, L.F64, 3, stats=False)
    got["inv"], _, e3 = _run(mcm, tape.trace(lambda p: p ** -1, 1), [x], n, L.F64, 3, stats=False)
    got["cdiv"], _, e4 = _run(mcm, tape.trace(lambda q: 3.0 / q, 1), [y], n, L.F64, 3, stats=False)
    got["divc"], _, e5 = _run(mcm, tape.trace(lambda p: p / 7.0, 1), [x], n, L.F64, 3, stats=False)
    assert (e1, e2, e3, e4, e5) == (3, 3, 3, 3, 3)
    for k_, w in want.items():
        g = got[k_][0]
        bad = ~((g == w) | (np.isnan(g) & np.isnan(w))) | (np.signbit(g) != np.signbit(w)) & ~np.isnan(w)
        assert not bad.any(), (k_, int(bad.sum()), x[bad][:4], y[bad][:4], g[bad][:4], w[bad][:4])
    # half-integer powers and sin / cos / exp over the whole range: the generated kernel equals the interpreter bit for bit
    ax = np.abs(x)
    for f in (lambda p: p ** 1.5, lambda p: p ** 2.5, lambda p: p ** -1.5, lambda p: p ** 0.5, lambda p: p ** 8.5,
              lambda p: mcm.sin(p) + mcm.cos(p), lambda p: mcm.exp(p)):
        t = tape.trace(f, 1)
        for arr in (ax, x):
            o3, _, e3 = _run(mcm, t, [arr], n, L.F64, 3, stats=False)
            o2, _, e2 = _run(mcm, t, [arr], n, L.F64, 2, stats=False)
            assert e3 == 3 and e2 == 2 and bits_equal(o3[0], o2[0])
    b = C.c_uint32(); L.check(L.lib().mcm_check_flags(C.byref(b), 1))
    # powers whose true value overflows are Inf, not NaN (the compensated product must not be used there)
    big = mcm.Particles(np.array([1e205, 1e250, 1e300, 1e-300, 1e100]))
    r = (big ** 1.5).to_numpy()
    assert r[0] > 1e307 and np.isinf(r[1]) and np.isinf(r[2]) and r[3] == 0.0 and abs(r[4] / 1e150 - 1) < 1e-15
    assert np.isinf((big ** 3.5).to_numpy()[4])
    L.check(L.lib().mcm_check_flags(C.byref(b), 1))


def test_elementwise_fast_forms_equal_the_exact_functions(mcm):
    """mcm_unary runs sin / cos / exp / sqrt / inv of Float64 vectors through branch-free fast forms with an exact
    redo outside their windows (csrc/elementwise_unary.cu k_map_fast): same bits as the per-value functions the
    interpreter uses, over the whole double range, for any length / alignment, and DomainError still raised."""
    from mcm_b200 import tape
    L = mcm._lib
    x = _wide_doubles(1 << 18, 5)
    with np.errstate(all="ignore"):
        want_sqrt, want_inv = np.sqrt(x), 1.0 / x
    for n in (len(x), 4097, 4099):
        xs = x[:n]
        p = mcm.Particles(xs)
        for f in (mcm.sin, mcm.cos, mcm.exp, mcm.sqrt, lambda q: q ** -1):
            got = f(p).to_numpy()
            ref, _, e2 = _run(mcm, tape.trace(f, 1), [xs], n, L.F64, 2, stats=False)
            assert e2 == 2 and bits_equal(got, ref[0])
        assert bits_equal(mcm.sqrt(p).to_numpy(), want_sqrt[:n]) and bits_equal((p ** -1).to_numpy(), want_inv[:n])
    b = C.c_uint32(); L.check(L.lib().mcm_check_flags(C.byref(b), 1))
    mcm.sqrt(mcm.Particles(np.concatenate([np.full(8192, 2.0), [-1.0]])))
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()
    mcm.sin(mcm.Particles(np.full(8192, 1e300)))           # huge but finite: library path, no DomainError
    mcm.check_domain()


def test_element_type_first_constructors(mcm):
    """runtests.jl:236-241: PT(Float32), PT(Float64), PT(Float32, 250), PT(Float32, 250, Normal(0.1f0)) and the
    ArgumentError for PT(Float32, 250, Gamma(0.1)) (types.jl:85-88: eltype(d) must equal T)."""
    a = mcm.Particles(np.float32)
    assert a.np_dtype == np.float32 and a.N == 2000
    assert mcm.Particles(np.float64).np_dtype == np.float64 and mcm.Particles(float).N == 2000
    b = mcm.Particles(np.float32, 250)
    assert b.np_dtype == np.float32 and b.N == 250
    c = mcm.Particles(np.float32, 250, mcm.Normal(np.float32(0.1)))
    assert c.np_dtype == np.float32 and c.N == 250 and abs(mcm.pmean(c) - 0.1) < 1e-6
    with pytest.raises(mcm.ArgumentError):
        mcm.Particles(np.float32, 250, mcm.Gamma(0.1))
    with pytest.raises(mcm.ArgumentError):
        mcm.Particles(np.float64, 250, mcm.Normal(np.float32(0.1)))
    g = mcm.Particles(np.float64, 300, mcm.Gamma(0.5), seed=2)
    assert g.N == 300 and g.np_dtype == np.float64 and g.to_numpy().min() > 0


def test_piterate_walks_the_particle_vector(mcm):
    """piterate(p) = iterate(p.particles) (particles.jl:316-323): Julia's (value, state) protocol."""
    x = np.array([3.0, -1.5, 8.25])
    p = mcm.Particles(x)
    got, st = [], None
    while True:
        r = mcm.piterate(p) if st is None else mcm.piterate(p, st)
        if r is None:
            break
        got.append(r[0]); st = r[1]
    assert got == list(x) and mcm.piterate(2.5) == (2.5, None) and mcm.piterate(2.5, 1) is None


def test_sigmapoints_like_the_reference(mcm):
    """runtests.jl:324-339: 2n+1 sigma points reproduce mean and covariance exactly (to rounding)."""
    s = mcm.sigmapoints(1, 3)
    assert s.shape == (3, 1) and abs(s.var(ddof=1) - 3) < 1e-14 and s.mean() == 1
    assert np.array_equal(mcm.sigmapoints(mcm.Normal(1, np.sqrt(3.0))), mcm.sigmapoints(1, np.sqrt(3.0) ** 2))
    m, S = [1, 2], np.array([[3.0, 1.0], [1.0, 4.0]])
    s = mcm.sigmapoints(m, S)
    assert s.shape == (5, 2) and np.allclose(np.cov(s.T), S, rtol=1e-13) and np.allclose(s.mean(axis=0), m, atol=1e-15)
    assert np.array_equal(mcm.sigmapoints(mcm.MvNormal(m, S)), s)
    p = mcm.Particles(s)                                      # StaticParticles(sigmapoints(m, Σ)) (sigmapoints.jl docstring)
    assert len(p) == 2 and p[0].N == 5 and np.allclose(mcm.pcov(p), S, rtol=1e-13) and np.allclose(mcm.pmean(p), m, atol=1e-15)


def test_moments_enqueue_collect_equals_the_synchronous_call(mcm):
    L = mcm._lib
    lib = L.lib()
    rng = np.random.default_rng(2)
    for n in (1, 1000, 3_000_001):
        x = mcm.Particles(rng.normal(5.0, 2.0, n))
        y = mcm.Particles(rng.normal(-1.0, 0.5, n))
        ref = np.zeros(6); got = np.zeros(6)
        L.check(lib.mcm_moments(L.F64, C.c_void_p(y.ptr), n, ref.ctypes.data_as(L._dp)))
        L.check(lib.mcm_moments_enqueue(L.F64, C.c_void_p(x.ptr), n))
        L.check(lib.mcm_moments_enqueue(L.F64, C.c_void_p(y.ptr), n))          # the last enqueue is the one collected
        L.check(lib.mcm_moments_collect(got.ctypes.data_as(L._dp)))
        assert np.array_equal(got, ref)


def test_has_particles_and_mean_object(mcm):
    """deconstruct.jl:5-20, 93-95 (runtests.jl uses them on ControlSystems-style structs): traversal on the host,
    means on the device."""
    import dataclasses

    @dataclasses.dataclass
    class Sys:
        A: list
        gain: object
        name: str

    p, q = mcm.pm(2.0, 0.1), mcm.pm(-1.0, 0.2)
    s = Sys(A=[[p, 1.0], [0.0, q]], gain=q, name="plant")
    assert mcm.has_particles(s) and mcm.has_particles([1.0, (2.0, p)]) and mcm.has_particles({"k": p})
    assert not mcm.has_particles(Sys(A=[[1.0]], gain=3.0, name="x")) and not mcm.has_particles(np.ones(3))
    m = mcm.mean_object(s)
    assert isinstance(m, Sys) and m.name == "plant" and not mcm.has_particles(m)
    assert m.A[0][0] == mcm.pmean(p) and m.A[1][1] == mcm.pmean(q) and m.gain == mcm.pmean(q) and m.A[0][1] == 1.0
    first = mcm.replace_particles({"a": p, "b": 4})
    assert first["a"] == p.to_numpy()[0] and first["b"] == 4
