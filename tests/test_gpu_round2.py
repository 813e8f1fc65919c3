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
