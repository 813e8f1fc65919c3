"""GPU tests for the §8(f) rows built on the hot path: optimize (optimize.jl:3-30), wasserstein
(distances.jl:7-13), ess (sampling.jl:26-38) and the iid constructors they need."""
import numpy as np
import pytest

from conftest import bits_equal, ulp_err

pytestmark = pytest.mark.gpu


def test_rand_real_matches_the_restated_stream(mcm, orc):
    n = 100003
    u = mcm.rand_particles(n, mcm.Uniform(0.0, 1.0), seed=11, offset=5).to_numpy()
    uo = orc.rand_uniform(n, 11, 5)
    assert bits_equal(u, uo) and u.min() > 0.0 and u.max() < 1.0
    z = mcm.rand_particles(n, mcm.Normal(1.5, 2.0), seed=11, offset=5)
    L = mcm._lib
    import ctypes as C
    pu = mcm.Particles(uo); ref = mcm.particles._empty(n, pu.dtype)
    L.check(L.lib().mcm_norminvcdf(pu.dtype, C.c_void_p(pu.ptr), C.c_void_p(ref.ptr), n))
    assert bits_equal(z.to_numpy(), 1.5 + 2.0 * ref.to_numpy())         # same device Phi^-1 on the same uniforms
    assert abs(z.to_numpy().mean() - 1.5) < 5 * 2 / np.sqrt(n) and abs(z.to_numpy().std() - 2.0) < 0.03
    a = mcm.rand_particles(1000, seed=1).to_numpy(); b = mcm.rand_particles(1000, seed=2).to_numpy()
    assert not np.array_equal(a, b) and np.array_equal(a, mcm.rand_particles(1000, seed=1).to_numpy())
    z32 = mcm.rand_particles(1000, seed=3, dtype=np.float32)
    assert z32.to_numpy().dtype == np.float32


@pytest.mark.parametrize("n", [4, 501, 10000, 300001])
def test_autocor_and_ess_match_oracle(mcm, orc, n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n)
    x = np.repeat(x, 2)[:n] if n > 100 else x                             # runtests.jl:32-34: duplicated neighbours
    p = mcm.Particles(x)
    import ctypes as C
    L = mcm._lib
    nl = min(250, n // 2)
    ac = np.zeros(nl)
    L.check(L.lib().mcm_autocor(p.dtype, C.c_void_p(p.ptr), n, nl, ac.ctypes.data_as(L._dp)))
    assert np.max(np.abs(ac - orc.autocor(x, nl))) <= 1e-12
    assert abs(mcm.ess(p) - orc.ess(x)) <= 1e-9 * n
    if n >= 10000:
        assert 0.45 * n < mcm.ess(p) < 0.56 * n                           # runtests.jl:34 window, scaled
        assert mcm.ess(mcm.Particles(rng.normal(size=n))) > 0.7 * n       # runtests.jl:31


def test_wasserstein(mcm, orc):
    p = mcm.pm(2.0, 0.1); q = mcm.pm(3.0, 0.1, seed=4)
    assert mcm.wasserstein(p, p, 1) == 0                                  # runtests.jl:555
    w1 = mcm.wasserstein(p, q, 1)
    a, b = np.sort(p.to_numpy()), np.sort(q.to_numpy())
    assert w1 >= 0 and abs(w1 - np.mean(np.abs(a - b))) <= 1e-14          # runtests.jl:556
    assert abs(mcm.wasserstein(p, q, 2) - np.mean(np.abs(a - b) ** 2)) <= 1e-13
    with pytest.raises(mcm.MethodError):
        mcm.wasserstein(p, mcm.Particles(np.zeros(5)), 1)


def test_optimize_rosenbrock(mcm):
    """runtests.jl:715-731: the cloud contracts onto the optimum [1, 1]; same seed => same result."""
    def rosenbrock2d(x):
        return (1.0 - x[0]) ** 2 + 100.0 * (x[1] - x[0] ** 2) ** 2
    ok = False
    for s in range(5):
        p = [-1.0 + 2.0 * mcm.Particles(200, mcm.Normal(), seed=10 * s + j) for j in range(2)]
        info = {}
        popt = mcm.optimize(rosenbrock2d, p, seed=s, info=info)
        m = np.array(mcm.pmean(popt))
        if np.all(np.abs(m - 1.0) < 0.05):
            ok = True
            break
    assert ok, (m, info)
    p = [-1.0 + 2.0 * mcm.Particles(200, mcm.Normal(), seed=j) for j in range(2)]
    r1 = mcm.optimize(rosenbrock2d, p, seed=876, iters=50)
    r2 = mcm.optimize(rosenbrock2d, p, seed=876, iters=50)
    assert all(bits_equal(a.to_numpy(), b.to_numpy()) for a, b in zip(r1, r2))
    assert p[0].N == 200                                                   # the input cloud is left alone


def test_perturb_has_the_requested_moments(mcm):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3)); Cp = A @ A.T + np.eye(3)
    p = [mcm.Particles(20000, mcm.Normal(float(j), 1.0), seed=j) for j in range(3)]
    q = mcm.perturb(p, Cp, seed=5)
    assert np.max(np.abs(np.array(mcm.pmean(q)) - np.array(mcm.pmean(p)))) < 0.05
    assert np.max(np.abs(mcm.pcov(q) - 1.1 * Cp)) < 0.05 * np.abs(Cp).max()


@pytest.mark.parametrize("n", [1, 7, 1000, 200003, 3_000_001])
def test_systematic_resampling_expansion_is_exact(mcm, orc, n):
    """Weighted systematic resampling (the weighted `sample(...; ordered=true)` + gather of optimize.jl:22-23,
    resampling.jl:45-67): identical indices from the expansion form, the per-output search and the oracle, for flat,
    zero-containing and extremely peaked weights (one source owning almost every output)."""
    lib = mcm._lib.lib()
    rng = np.random.default_rng(n)
    cases = {
        "flat": np.ones(n),
        "lognormal": np.exp(3 * rng.normal(size=n)),
        "zeros": np.where(rng.random(n) < 0.7, 0.0, rng.random(n)) + (np.arange(n) == n // 2),
        "peaked": np.exp(-0.5 * ((np.arange(n) - 0.3 * n) / max(1.0, 1e-4 * n)) ** 2) + 1e-300,
        "one": (np.arange(n) == n - 1).astype(np.float64),
    }
    for name, w in cases.items():
        for u in (0.0, 0.5, 0.999999):
            want = orc.resample_systematic(w, u)
            got = mcm.resample_indices(w, u).to_numpy()
            assert np.array_equal(got, want), (name, u)
        lib.mcm_resample_use_search(1)
        try:
            assert np.array_equal(mcm.resample_indices(w, 0.25).to_numpy(), orc.resample_systematic(w, 0.25)), name
        finally:
            lib.mcm_resample_use_search(0)


def test_sharded_bootstrap_single_rank_equals_bootstrap(mcm):
    """world = 1: the peer-window path (window = own memory) reproduces bootstrap(p, n, seed) bit for bit."""
    rng = np.random.default_rng(2)
    X = rng.normal(size=(5, 100_001))
    ps = [mcm.Particles(np.ascontiguousarray(x)) for x in X]
    for n in (None, 40_000):
        a = mcm.sharded.bootstrap(ps, X.shape[1], n, seed=11, force_window=True)
        b = mcm.bootstrap(ps, n, seed=11)
        assert all(np.array_equal(p.to_numpy(), q.to_numpy()) for p, q in zip(a, b))
    w = rng.random(X.shape[1]) ** 3
    for u in (0.25, 0.0):
        a = mcm.sharded.resample_indices(w, X.shape[1], u).to_numpy()
        assert np.array_equal(a, mcm.resample_indices(w, u).to_numpy())
    mcm.sharded.free_windows()
    lib = mcm._lib.lib()
    assert lib.mcm_window_free(12345) == mcm._lib.MCM_ERR_INVALID
    assert lib.mcm_comm_fence() == 0


def test_optimize_trajectory_step_by_step_against_the_oracle(mcm, orc):
    """optimize.jl:14-30 replayed on the host with the SAME Philox streams: every iteration's weights, ancestors
    (bit-exact integer scan), gathered cloud (bit-exact), Cp (1e-12) and re-perturbed cloud
    (mu + chol(1.1 Cp + 1e-12 I) z, z = Phi^-1 of the Philox uniforms keyed by (seed, j N)) are checked against the
    oracle from that iteration's own inputs, and the loop wiring (tau schedule, one uniform offset and one seed per
    iteration from the host generator) against a host restatement of the loop."""
    def rosenbrock2d(x):
        return (1.0 - x[0]) ** 2 + 100.0 * (x[1] - x[0] ** 2) ** 2
    N, iters, seed0 = 500, 12, 321
    p0 = [-1.0 + 2.0 * mcm.Particles(N, mcm.Normal(), seed=j) for j in range(2)]
    trace = []
    mcm.optimize(rosenbrock2d, p0, tau=1.0, tau_i=1.005, iters=iters, tol=0.0, seed=seed0, trace=trace)
    assert len(trace) == iters
    rng = np.random.default_rng(seed0)
    cloud = [x.to_numpy() for x in p0]
    tau = 1.0
    for it, tr in enumerate(trace):
        assert tr["tau"] == tau
        y = -((1.0 - cloud[0]) * (1.0 - cloud[0]) + 100.0 * ((cloud[1] - cloud[0] * cloud[0]) * (cloud[1] - cloud[0] * cloud[0])))
        assert bits_equal(tr["y"], y), it                                   # + - * and literal powers are bit-exact
        assert ulp_err(tr["we"], orc.unary("EXP", tau * y)).max() <= 2
        u = float(rng.random())
        assert tr["u"] == u
        j = orc.resample_systematic(tr["we"], u)                           # from the device's own weights
        assert np.array_equal(tr["j"], j), it
        for k in range(2):
            assert bits_equal(tr["gathered"][k], cloud[k][j])
        G = np.stack(tr["gathered"])
        Co = orc.cov(np.ascontiguousarray(G))
        assert np.max(np.abs(tr["Cp"] - Co)) <= 1e-12 * np.max(np.abs(Co)), it
        pseed = int(rng.integers(0, 2 ** 62))
        assert tr["seed"] == pseed
        Lc = np.linalg.cholesky(1.1 * tr["Cp"] + 1e-12 * np.eye(2))
        z = np.stack([orc.norminvcdf(orc.rand_uniform(N, pseed, k * N)) for k in range(2)])
        mu = np.array([orc.moments(G[k])[1] for k in range(2)])
        want = mu[:, None] + Lc @ z
        got = np.stack(tr["perturbed"])
        scale = np.abs(Lc).max() * np.abs(z).max() + np.abs(mu).max()
        assert np.max(np.abs(got - want)) <= 1e-11 * scale, (it, np.max(np.abs(got - want)))
        cloud = [got[0], got[1]]                                           # continue from the device's cloud
        tau *= 1.005
