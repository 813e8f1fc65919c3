"""union / intersect / with_nominal / nominal on device vs the oracle restatement
(reference: src/particles.jl:360-381, src/nominal.jl; reference tests: test/runtests.jl `union`/`intersect`
and test/test_nominal.jl-style checks: nominal(with_nominal(p, v)) == v, one particle replaced)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_union_is_concatenation(mcm, orc, dt):
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=2000).astype(dt), rng.normal(1, 2, size=777).astype(dt)
    u = mcm.union(mcm.Particles(a), mcm.Particles(b))
    assert u.N == 2777
    assert np.array_equal(u.to_numpy(), orc.union(a, b))
    with pytest.raises(mcm.MethodError):
        mcm.union(mcm.Particles(a.astype(np.float64)), mcm.Particles(b.astype(np.float32)))


@pytest.mark.parametrize("n1,n2", [(2000, 2000), (100_003, 65_537), (5, 3_000_001)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_intersect_matches_filter(mcm, orc, n1, n2, dt):
    rng = np.random.default_rng(n1 + n2)
    a = rng.normal(0, 1, size=n1).astype(dt)
    b = rng.normal(0.5, 0.7, size=n2).astype(dt)
    got = mcm.intersect(mcm.Particles(a), mcm.Particles(b)).to_numpy()
    want = orc.intersect(a, b)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    # every survivor is inside the common support, and the ends of the support are kept
    assert got.min() == max(a.min(), b.min()) and got.max() == min(a.max(), b.max())


def test_intersect_edge_cases(mcm, orc):
    # identical supports keep everything; ties at the ends are inclusive
    a = np.array([0.0, 1.0, 2.0, 2.0, 0.0, 1.5])
    got = mcm.intersect(mcm.Particles(a), mcm.Particles(a[::-1].copy())).to_numpy()
    assert np.array_equal(got, np.concatenate([a, a[::-1]]))
    # disjoint supports: an empty Particles, as the reference builds
    e = mcm.intersect(mcm.Particles(np.array([0.0, 1.0])), mcm.Particles(np.array([2.0, 3.0])))
    assert e.N == 0 and e.to_numpy().size == 0
    # statistics of the result work on the logical length
    x, y = np.linspace(0, 1, 1001), np.linspace(0.5, 1.5, 1001)
    p = mcm.intersect(mcm.Particles(x), mcm.Particles(y))
    ref = orc.intersect(x, y)
    assert p.N == ref.size and abs(mcm.pmean(p) - ref.mean()) < 1e-14


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [2000, 1_000_003])
def test_with_nominal_scalar(mcm, orc, dt, n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n).astype(dt)
    for val in (0.25, -3.0, 100.0):
        p = mcm.with_nominal(mcm.Particles(x), val)
        want = orc.with_nominal(x, dt(val))
        assert np.array_equal(p.to_numpy(), want)
        assert mcm.nominal(p) == dt(val)
        # exactly one particle of the original went missing (the closest one), unless it was already at index 0
        assert np.sum(p.to_numpy() != x) <= 2


def test_with_nominal_ties_and_nan(mcm, orc):
    x = np.array([1.0, 3.0, -1.0, 3.0, 1.0])
    # |x - 2| is minimal (1.0) at indices 0, 1, 3, 4: the first wins, as Julia's argmin
    p = mcm.with_nominal(mcm.Particles(x), 2.0)
    assert np.array_equal(p.to_numpy(), orc.with_nominal(x, 2.0)) and p.to_numpy()[0] == 2.0
    y = np.array([1.0, 2.0, np.nan, 0.5, np.nan])
    got = mcm.with_nominal(mcm.Particles(y), 0.5).to_numpy()
    want = orc.with_nominal(y, 0.5)       # the first NaN is the argmin
    assert np.array_equal(got, want, equal_nan=True) and got[2] == 1.0 and got[0] == 0.5


@pytest.mark.parametrize("d,n", [(3, 2000), (70, 50_000)])
def test_with_nominal_vector(mcm, orc, d, n):
    rng = np.random.default_rng(d)
    X = rng.normal(size=(n, d)) * np.linspace(0.5, 2.0, d)
    vals = rng.normal(size=d) * 0.1
    ps = [mcm.Particles(np.ascontiguousarray(X[:, j])) for j in range(d)]
    # joint nearest particle
    got = mcm.with_nominal(ps, vals, joint=True)
    want, k = orc.with_nominal_mv(X, vals, joint=True)
    for g, w in zip(got, want):
        assert np.array_equal(g.to_numpy(), w)
    assert mcm.nominal(got) == list(vals)
    # literal reference behaviour (column sums, index reused as a particle index)
    got = mcm.with_nominal(ps, vals)
    want, k = orc.with_nominal_mv(X, vals, joint=False)
    assert k == 0          # the column with the smallest scale
    for g, w in zip(got, want):
        assert np.array_equal(g.to_numpy(), w)


def test_decision_entry_points_reject_bad_arguments(mcm):
    import ctypes as C
    L = mcm._lib
    lib = L.lib()
    idx = C.c_uint64(0)
    assert lib.mcm_argmin_absdiff(L.F64, None, 10, 0.0, C.byref(idx)) == L.MCM_ERR_INVALID
    p = mcm.Particles(np.arange(4.0))
    assert lib.mcm_shuffle_insert(L.F64, C.c_void_p(p.ptr), 4, 4, 0.0, C.c_void_p(p.ptr)) == L.MCM_ERR_INVALID
    assert lib.mcm_argmin_absdiff(99, C.c_void_p(p.ptr), 4, 0.0, C.byref(idx)) == L.MCM_ERR_INVALID
    k = C.c_size_t(7)
    assert lib.mcm_filter_range(L.F64, None, 0, 0.0, 1.0, None, 0, C.byref(k)) == 0 and k.value == 0


def test_reference_testset_union_intersect_nominal(mcm):
    """test/runtests.jl:525-527 and :930-956, statement by statement."""
    p = mcm.pm(1.0, 0.1)
    assert np.array_equal(mcm.intersect(p, p).to_numpy(), mcm.union(p, p).to_numpy())
    assert mcm.intersect(p, 1 + p).N < 2 * p.N
    assert mcm.union(p, 1 + p).N == 2 * p.N
    for make in (lambda: mcm.pm(1.0, 0.1), lambda: mcm.mp(1.0, 0.1)):
        p = make()
        pn = mcm.with_nominal(p, 0)
        assert mcm.nominal(pn) == pn.to_numpy()[0] == 0
        assert mcm.nominal(p) != 0
        pv = [p, p]
        pn = mcm.with_nominal(pv, [0, 1])
        assert mcm.nominal(pn) == [q.to_numpy()[0] for q in pn] == [0, 1]
        assert mcm.nominal(pv) != [0, 1]
