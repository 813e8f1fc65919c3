"""Full-size checks (BASELINE.json sizes) through size-independent properties — the oracle cannot
finish N = 1e8 in seconds, so these use sortedness, checksums, round trips, idempotence, agreement of
independent device paths, and spot slices against the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

N8 = 100_000_000


def _count_unsorted(mcm, p):
    """number of i with p[i+1] < p[i] — adjacent differences through the C ABI on an offset view"""
    L = mcm._lib
    n = p.N - 1
    d = mcm.particles._empty(n, p.dtype)
    esz = 8 if p.dtype == L.F64 else 4
    L.check(L.lib().mcm_binary(L.ENUMS["MCM_B_SUB"], p.dtype, C.c_void_p(p.ptr + esz), 0.0, C.c_void_p(p.ptr), 0.0,
                               C.c_void_p(d.ptr), n))
    return mcm.particles._count("LT", d, 0.0)


def _slice_to_numpy(mcm, p, lo, hi):
    L = mcm._lib
    out = np.empty(hi - lo, dtype=p.np_dtype)
    L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(p.ptr + lo * out.itemsize), out.nbytes))
    return out


@pytest.fixture(scope="module")
def a8(mcm):
    return mcm.systematic_sample(N8, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)


def test_c2_systematic_sample_1e8(mcm, orc, a8):
    assert _count_unsorted(mcm, a8) == 0                         # permute=false -> sorted
    m = mcm.moments(a8)
    assert m.n == N8 and abs(m.mean - np.pi) <= 1e-13            # symmetric grid
    from scipy.special import ndtri
    z = ndtri((2.0 * np.arange(N8 // 50) + 1) / (2.0 * N8))      # the lower 2 % tail on the host
    head = _slice_to_numpy(mcm, a8, 0, N8 // 50)
    assert np.max(np.abs(head - (np.pi + 0.1 * z))) <= 0.1 * 6e-15 * 6 + 1e-15
    # spot slices against the oracle (first, middle, last 1000 grid points)
    for lo in (0, N8 // 2 - 500, N8 - 1000):
        ref = orc.systematic_normal(N8, np.pi, 0.1, i0=lo, n_total=N8, n_local=1000)
        got = _slice_to_numpy(mcm, a8, lo, lo + 1000)
        assert np.max(np.abs(got - ref)) <= 0.1 * 6 * 2.3e-16 * 8 + 4.5e-16
    # antisymmetry about the mean, limited by the conditioning of the upper tail: p = 1 - 5e-9 carries an
    # absolute rounding error of 1.1e-16, amplified by 1/phi(z) ~ 3e7 (the reference has the same property)
    tail = _slice_to_numpy(mcm, a8, N8 - 1000, N8)[::-1]
    assert np.max(np.abs((head[:1000] - np.pi) + (tail - np.pi))) <= 1e-9


def test_c2_chain10_1e8_paths_agree(mcm, orc, a8):
    from mcm_b200 import chains
    f = mcm.fused(chains.chain10)
    y, st = f(a8, stats=True)
    _, st2 = f(a8, stats=True, materialize=False)                 # statistics-only launch: 8 B/particle
    assert (st.n, st.mean, st.m2, st.min, st.max) == (st2.n, st2.mean, st2.m2, st2.min, st2.max)
    m = mcm.moments(y)                                            # independent kernel over the materialised y
    assert m.n == N8 and abs(m.mean - st.mean) <= 1e-13 * abs(st.mean) and abs(m.m2 - st.m2) <= 1e-12 * st.m2
    assert m.min == st.min and m.max == st.max
    for lo in (0, 33_333_333, N8 - 4096):                          # spot slices against the oracle
        ah = _slice_to_numpy(mcm, a8, lo, lo + 4096)
        yh = _slice_to_numpy(mcm, y, lo, lo + 4096)
        assert np.all(np.abs(yh - orc.chain10(ah)) <= orc.chain10_bound(ah))
    # unfused evaluation (one launch per primitive) is bit-identical on a 1e7 prefix
    n = 10_000_000
    ap = mcm.Particles(_slice_to_numpy(mcm, a8, 0, n))
    y3 = mcm.sqrt(abs(mcm.sin(ap) * ap ** 2 + mcm.exp(-ap))) / ap ** 1.5
    assert bits_equal(y3.to_numpy(), _slice_to_numpy(mcm, y, 0, n))


def test_sort_quantile_1e8(mcm, a8):
    from mcm_b200 import chains
    y = mcm.fused(chains.chain10)(a8)                             # non-monotone function of a sorted input
    s = mcm.psort(y)
    assert _count_unsorted(mcm, s) == 0
    my, ms = mcm.moments(y), mcm.moments(s)
    assert (ms.n, ms.min, ms.max) == (my.n, my.min, my.max)       # same multiset: checksums agree
    assert abs(ms.mean - my.mean) <= 1e-13 * abs(my.mean) and abs(ms.m2 - my.m2) <= 1e-12 * my.m2
    assert _count_unsorted(mcm, mcm.psort(s)) == 0 and mcm.particles._count("NE", mcm.psort(s) - s, 0.0) == 0  # idempotent
    q = mcm.pquantile(y, [0.0, 0.5, 1.0])
    assert q[0] == my.min and q[2] == my.max
    mid = _slice_to_numpy(mcm, s, N8 // 2 - 1, N8 // 2 + 1)
    assert q[1] == mid[0] + ((N8 * 0.5 + 0.5) - N8 // 2) * (mid[1] - mid[0])
    assert mcm.pmedian(a8) == (lambda v: v[0] / 2 + v[1] / 2)(_slice_to_numpy(mcm, a8, N8 // 2 - 1, N8 // 2 + 1))


def test_permutation_and_bootstrap_round_trips_1e7(mcm):
    L = mcm._lib
    n = 10_000_000
    x = mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=False, as_particles=True)
    perm = mcm.DeviceIndex(n)
    L.check(L.lib().mcm_randperm(C.c_void_p(perm.ptr), n, 12345))
    xp = mcm.bootstrap(x, n, inds=perm)                           # x[perm]: a permutation of a sorted vector
    assert _count_unsorted(mcm, xp) > n // 3                       # really shuffled
    assert mcm.particles._count("NE", mcm.psort(xp) - x, 0.0) == 0  # sort undoes it exactly => perm is a bijection
    sp = mcm.psortperm(xp)                                          # and sortperm gives the inverse route
    assert mcm.particles._count("NE", mcm.bootstrap(xp, n, inds=sp) - x, 0.0) == 0
    ident = mcm.particles._upload_i64(np.arange(n, dtype=np.int64))
    assert mcm.particles._count("NE", mcm.bootstrap(x, n, inds=ident) - x, 0.0) == 0
    # uniform bootstrap: index histogram is flat; resampled mean within Monte-Carlo error
    b = mcm.bootstrap(x, seed=9)
    assert abs(mcm.pmean(b)) < 5 / np.sqrt(n) and abs(mcm.pstd(b) - 1) < 5 / np.sqrt(n)
    # systematic resampling with equal weights is the identity; with weights ~ density ratio it shifts the mean
    w = mcm.fill(1.0, n)
    assert mcm.particles._count("NE", mcm.resample(x, w, 0.25) - x, 0.0) == 0
    r = mcm.resample(x, mcm.exp(0.5 * x), 0.5)                     # tilting N(0,1) by exp(x/2) -> N(0.5, 1)
    assert abs(mcm.pmean(r) - 0.5) < 1e-3 and abs(mcm.pstd(r) - 1.0) < 1e-3


def test_c4_pcov_4096_by_1e5(mcm, orc):
    M, n = 4096, 100_000
    L = mcm._lib
    base = [mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=True, seed=j + 1, as_particles=True) for j in range(64)]
    v = []
    for j in range(M):                                            # banded mixing of 64 independent columns
        b0, b1 = base[j % 64], base[(j * 7 + 3) % 64]
        v.append(mcm.lincomb([1.0 + 0.001 * j, 0.5], [b0, b1]) if j % 64 != (j * 7 + 3) % 64 else b0 * (1.0 + 0.001 * j))
    v[100] = v[5]                                                 # an exact duplicate column
    Cg = mcm.pcov(v)
    assert Cg.shape == (M, M) and np.array_equal(Cg, Cg.T)
    d = np.array([mcm.pvar(v[j]) for j in (0, 5, 100, 777, 4095)])
    assert np.max(np.abs(np.diag(Cg)[[0, 5, 100, 777, 4095]] - d)) <= 1e-12 * d.max()
    assert Cg[100, 5] == Cg[5, 5] == Cg[100, 100]
    sub = [0, 5, 64, 100, 1000, 2047, 2048, 4095]
    X = np.stack([v[j].to_numpy() for j in sub])
    Cr = orc.cov(X)
    assert np.max(np.abs(Cg[np.ix_(sub, sub)] - Cr)) <= 1e-12 * np.max(np.abs(Cr))


def test_c3_pendulum_1e7_fused_vs_stepwise(mcm):
    from mcm_b200 import chains, tape
    n, steps = 10_000_000, 50
    L = mcm._lib
    mk = lambda mu, s, seed: mcm.systematic_sample(n, mcm.Normal(mu, s), permute=True, seed=seed, as_particles=True)
    th, om = mcm.fill(0.0, n), mk(np.pi / 3, 0.02, 1)
    g, Lp, c = mk(9.79, 0.02, 2), mk(1.0, 0.01, 3), mk(0.1, 0.01, 4)
    t = chains.pendulum_rk4_tape(0.01)
    o1, o2 = mcm.particles._empty(n, L.F64), mcm.particles._empty(n, L.F64)
    tape.run_tape(t, [p.ptr for p in (th, om, g, Lp, c)], n, L.F64, [o1.ptr, o2.ptr], reps=steps)
    # the same integration, one launch per RK4 step (trajectory-saving variant): bit-identical
    s1, s2 = th, om
    for _ in range(steps):
        n1, n2 = mcm.particles._empty(n, L.F64), mcm.particles._empty(n, L.F64)
        tape.run_tape(t, [s1.ptr, s2.ptr, g.ptr, Lp.ptr, c.ptr], n, L.F64, [n1.ptr, n2.ptr], reps=1)
        s1, s2 = n1, n2
    assert mcm.particles._count("NE", s1 - o1, 0.0) == 0 and mcm.particles._count("NE", s2 - o2, 0.0) == 0
    # energy of the damped pendulum must not increase
    gL = g / Lp
    e0 = 0.5 * om * om + gL * (1.0 - mcm.cos(th))
    e1 = 0.5 * o2 * o2 + gL * (1.0 - mcm.cos(o1))
    assert mcm.particles._count("GT", e1 - e0, 1e-9) == 0
