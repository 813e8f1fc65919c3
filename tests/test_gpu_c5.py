"""GPU parity for the C5 pipeline pieces and Particle BLAS (§8f-1):
batched keyed systematic columns, constant-matrix application (mcm_pgemm), shared-index gather over many
variables, batched quantiles.  Index / order-statistic outputs bit-exact; pgemm within a few rounding errors of
the long-double oracle (tolerance stated per test)."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


# ---- keyed index bijection + batched systematic columns ------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 4097, 65536, 300001])
def test_index_bijection_matches_oracle_and_is_a_permutation(mcm, orc, n):
    for seed, col in ((0, 0), (7, 3), (2**63 + 5, 1023)):
        got = mcm.index_bijection(n, seed=seed, column=col).to_numpy()
        assert np.array_equal(got, orc.index_bijection(n, seed, col))
        assert np.array_equal(np.sort(got), np.arange(n))


def test_index_bijection_shards_compose(mcm, orc):
    n = 100003
    full = mcm.index_bijection(n, seed=11).to_numpy()
    cuts = [0, 1, 33333, 70000, n]
    parts = [mcm.index_bijection(b - a, seed=11, i0=a, n_total=n).to_numpy() for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), full)
    with pytest.raises(mcm.MCMError):
        mcm.index_bijection(10, i0=95, n_total=100)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_systematic_batch_equals_per_column_sampling(mcm, orc, dt):
    n, M = 20011, 5
    dists = [mcm.Normal(0.5 * j, 1.0 + j) for j in range(M)]
    cols = mcm.systematic_batch(n, dists, seed=9, dtype=dt)
    for j, (c, d) in enumerate(zip(cols, dists)):
        perm = orc.index_bijection(n, 9, j)
        dd = mcm.Normal(dt(d.mu), dt(d.sigma)) if dt == np.float32 else d
        ref = mcm.systematic_sample(n, dd, perm=perm, as_particles=True)       # same device Phi^-1, explicit permutation
        assert bits_equal(c.to_numpy(), ref.to_numpy()), j
    # unpermuted == grid order, Uniform family, standard parameters
    u = mcm.systematic_batch(n, [mcm.Uniform(-1.0, 3.0)], permute=False)[0].to_numpy()
    assert bits_equal(u, orc.systematic_uniform(n, -1.0, 3.0))
    z = mcm.systematic_batch(n, 2, seed=1)
    assert bits_equal(np.sort(z[0].to_numpy()), np.sort(z[1].to_numpy()))      # same multiset, different order
    assert not np.array_equal(z[0].to_numpy(), z[1].to_numpy())
    assert abs(np.corrcoef(z[0].to_numpy(), z[1].to_numpy())[0, 1]) < 5 / np.sqrt(n)
    # sharded generation composes
    a = mcm.systematic_batch(n, 1, seed=3, i0=0, n_total=n, n_local=8000)[0].to_numpy()
    b = mcm.systematic_batch(n, 1, seed=3, i0=8000, n_total=n, n_local=n - 8000)[0].to_numpy()
    assert bits_equal(np.concatenate([a, b]), mcm.systematic_batch(n, 1, seed=3)[0].to_numpy())


# ---- constant matrix x particle vector ---------------------------------------------------------
def _pgemm_bound(A, X, shift):
    """|error| <= (Mi + 4) eps * sum_k |a_jk| |x_k - s_k|: standard dot-product bound, any summation order."""
    Xa = np.abs(X - (0 if shift is None else np.asarray(shift)[:, None]))
    return (A.shape[1] + 4) * np.finfo(np.float64).eps * (np.abs(A) @ Xa)


@pytest.mark.parametrize("Mo,Mi,n", [(1, 1, 1), (3, 5, 1000), (128, 16, 128), (130, 37, 777), (257, 200, 1031)])
def test_pgemm_matches_long_double_oracle(mcm, orc, Mo, Mi, n):
    rng = np.random.default_rng(Mo * 7 + Mi)
    A = rng.normal(size=(Mo, Mi))
    X = rng.normal(2.0, 3.0, size=(Mi, n))
    shift = rng.normal(2.0, 0.1, size=Mi)
    bias = rng.normal(size=Mo)
    xs = [mcm.Particles(r) for r in X]
    Y = np.stack([y.to_numpy() for y in mcm.pgemm(A, xs, shift=shift, bias=bias)])
    ref = orc.pgemm(A, X, shift, bias)
    assert np.all(np.abs(Y - ref) <= _pgemm_bound(A, X, shift) + 1e-300)
    Y0 = np.stack([y.to_numpy() for y in mcm.pgemm(A, xs)])                      # no shift, no bias
    assert np.all(np.abs(Y0 - orc.pgemm(A, X)) <= _pgemm_bound(A, X, None) + 1e-300)


def test_pgemm_triangular_unaligned_float32_and_offset(mcm, orc):
    rng = np.random.default_rng(5)
    Mi = Mo = 150
    n = 2049
    A = np.tril(rng.normal(size=(Mo, Mi)))                                       # k loop stops early per row tile
    X = rng.normal(size=(Mi, n))
    xs = [mcm.Particles(r) for r in X]
    Y = np.stack([y.to_numpy() for y in mcm.pgemm(A, xs)])
    assert np.all(np.abs(Y - orc.pgemm(A, X)) <= _pgemm_bound(A, X, None))
    # exact small-integer data: every product and sum is exact in Float64 -> bit-equal to the oracle
    Ai = np.tril(rng.integers(-8, 9, size=(Mo, Mi))).astype(np.float64)
    Xi = rng.integers(-1000, 1001, size=(Mi, n)).astype(np.float64)
    Yi = np.stack([y.to_numpy() for y in mcm.pgemm(Ai, [mcm.Particles(r) for r in Xi], bias=np.arange(Mo))])
    assert np.array_equal(Yi, orc.pgemm(Ai, Xi, None, np.arange(Mo, dtype=np.float64)))
    # views at odd offsets (8-byte aligned only): 8-byte cp.async path, through the C ABI with raw pointers
    L = mcm._lib
    base = [mcm.Particles(np.concatenate([[0.0], r])) for r in X[:20]]
    outs = [mcm.particles._empty(n + 1, base[0].dtype) for _ in range(20)]
    pin = (C.c_void_p * 20)(*[C.c_void_p(b.ptr + 8) for b in base])
    pout = (C.c_void_p * 20)(*[C.c_void_p(o.ptr + 8) for o in outs])
    A20 = np.ascontiguousarray(A[:20, :20])
    L.check(L.lib().mcm_pgemm(base[0].dtype, A20.ctypes.data_as(L._dp), 20, 20, pin, None, None, pout, n))
    Yv = np.stack([o.to_numpy()[1:] for o in outs])
    assert np.all(np.abs(Yv - orc.pgemm(A20, X[:20])) <= _pgemm_bound(A20, X[:20], None))
    # Float32 particles: Float64 accumulation, rounded once on store
    X32 = X[:40].astype(np.float32)
    Y32 = np.stack([y.to_numpy() for y in mcm.pgemm(A[:40, :40], [mcm.Particles(r) for r in X32])])
    ref32 = orc.pgemm(A[:40, :40], X32.astype(np.float64))
    assert Y32.dtype == np.float32 and np.all(np.abs(Y32 - ref32) <= 6e-8 * np.abs(ref32) + 1e-5 * 1e-2)
    # centring at fragment load: data at 1e8 +- 1 keeps full relative accuracy of the centred values
    Xo = 1e8 + rng.normal(size=(8, 500))
    mu = Xo.mean(axis=1)
    Yo = np.stack([y.to_numpy() for y in mcm.pgemm(np.eye(8) * 3.0, [mcm.Particles(r) for r in Xo], shift=mu)])
    assert np.array_equal(Yo, 3.0 * (Xo - mu[:, None]))
    with pytest.raises(mcm.ArgumentError):
        mcm.pgemm(np.ones((2, 3)), xs[:2])


def _ozaki_floor(A, X, shift):
    """Absolute error floor of the 7-digit int8 split: operands are kept to 2^-56 of the row maximum
    max_k |A_jk| max_i |x_k - s_k| (x 4 for the two power-of-two roundings), K products per output."""
    Xc = np.abs(X - (0 if shift is None else np.asarray(shift)[:, None]))
    rowmax = np.max(np.abs(A) * Xc.max(axis=1)[None, :], axis=1)
    return (8 * A.shape[1] * 2.0 ** -56 * 4 * rowmax)[:, None]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("Mo,Mi,n,tri", [(64, 128, 8192, False), (130, 200, 10001, False), (257, 300, 20000, True)])
def test_pgemm_tcgen05_engine_vs_oracle_and_dmma(mcm, orc, dt, Mo, Mi, n, tri):
    """Large problems run on the tcgen05/TMEM int8 path (7 balanced 8-bit digits, exact int32 accumulation); both
    engines must agree with the long-double oracle, the tensor path additionally within its stated floor."""
    lib = mcm._lib.lib()
    rng = np.random.default_rng(Mo + Mi)
    A = rng.normal(size=(Mo, Mi)) * np.exp(rng.normal(size=(Mo, 1)))          # rows of different magnitude
    if tri:
        A = np.tril(A)
    X = (rng.normal(size=(Mi, n)) * np.exp(2 * rng.normal(size=(Mi, 1))) + rng.normal(size=(Mi, 1))).astype(dt)
    shift = X.astype(np.float64).mean(axis=1)
    bias = rng.normal(size=Mo)
    xs = [mcm.Particles(r) for r in X]
    Yt = np.stack([y.to_numpy() for y in mcm.pgemm(A, xs, shift=shift, bias=bias)])
    assert lib.mcm_pgemm_last_engine() == 1
    lib.mcm_pgemm_use_tensor(0)
    try:
        Yd = np.stack([y.to_numpy() for y in mcm.pgemm(A, xs, shift=shift, bias=bias)])
        assert lib.mcm_pgemm_last_engine() == 0
    finally:
        lib.mcm_pgemm_use_tensor(1)
    sub = sorted({0, 1, 63, min(64, Mo - 1), Mo // 2, Mo - 1})
    cols = np.r_[0:300, n - 300:n]
    X64 = X.astype(np.float64)
    ref = orc.pgemm(A[sub], X64[:, cols], shift, bias[sub])
    bound = _pgemm_bound(A[sub], X64[:, cols], shift)
    if dt == np.float32:
        bound = bound + 6e-8 * np.abs(ref)                                     # one rounding to Float32 on store
    assert np.all(np.abs(Yd[np.ix_(sub, cols)] - ref) <= bound + 1e-300)
    assert np.all(np.abs(Yt[np.ix_(sub, cols)] - ref) <= bound + _ozaki_floor(A[sub], X64, shift))
    # whole result: the two engines agree to the same floor
    full_floor = _ozaki_floor(A, X64, shift) + (Mi + 4) * 2.3e-16 * (np.abs(A) @ np.abs(X64 - shift[:, None]))
    if dt == np.float32:
        full_floor = full_floor + 1.2e-7 * np.abs(Yd)
    assert np.all(np.abs(Yt.astype(np.float64) - Yd) <= full_floor)


def test_pgemm_tcgen05_declines_what_it_cannot_represent(mcm, orc):
    lib = mcm._lib.lib()
    rng = np.random.default_rng(77)
    A = rng.normal(size=(64, 128))
    Xh = rng.standard_t(2, size=(128, 8192))                                   # heavy tails: max/sigma >> 40
    Y = np.stack([y.to_numpy() for y in mcm.pgemm(A, [mcm.Particles(r) for r in Xh])])
    assert lib.mcm_pgemm_last_engine() == 0
    assert np.all(np.abs(Y[:2, :200] - orc.pgemm(A[:2], Xh[:, :200])) <= _pgemm_bound(A[:2], Xh[:, :200], None))
    Xc = rng.normal(size=(128, 8192)); Xc[7] = 2.5                             # a constant variable is fine (digits 0)
    Yc = np.stack([y.to_numpy() for y in mcm.pgemm(A, [mcm.Particles(r) for r in Xc], shift=Xc.mean(axis=1))])
    assert lib.mcm_pgemm_last_engine() == 1
    refc = orc.pgemm(A[:3], Xc[:, :300], Xc.mean(axis=1), None)
    assert np.all(np.abs(Yc[:3, :300] - refc) <= _pgemm_bound(A[:3], Xc[:, :300], Xc.mean(axis=1)) + _ozaki_floor(A[:3], Xc, Xc.mean(axis=1)))
    Az = A.copy(); Az[5] = 0.0                                                 # an all-zero output row -> exactly the bias
    Yz = mcm.pgemm(Az, [mcm.Particles(r) for r in Xc], bias=np.full(64, 1.5))
    assert lib.mcm_pgemm_last_engine() == 1 and np.all(Yz[5].to_numpy() == 1.5)


def test_particle_blas_wrappers(mcm, orc):
    """mul!(y, A, b), A*p, dot(v, p), axpy!(a, x, y) (particles.jl:620-707; runtests.jl:735-753 pattern)."""
    rng = np.random.default_rng(8)
    A = rng.normal(size=(3, 4))
    X = rng.normal(size=(4, 500))
    b = [mcm.Particles(r) for r in X]
    naive = [sum(float(A[j, k]) * b[k] for k in range(4)) for j in range(3)]      # the "naive way" of the reference
    for y, z in zip(mcm.pgemv(A, b), naive):
        assert np.allclose(y.to_numpy(), z.to_numpy(), rtol=0, atol=1e-14 * 20)
    y = [None] * 3
    assert mcm.mul_(y, A, b) is y and bits_equal(y[1].to_numpy(), mcm.pgemv(A, b)[1].to_numpy())
    d = mcm.pdot(A[0], b)
    assert np.allclose(d.to_numpy(), naive[0].to_numpy(), rtol=0, atol=1e-14 * 20)
    x2 = [mcm.Particles(r) for r in X[:2]]; y2 = [mcm.Particles(r) for r in X[2:]]
    mcm.paxpy_(2.5, x2, y2)
    assert bits_equal(y2[0].to_numpy(), np.array([np.float64(2.5) * u + v for u, v in zip(X[0], X[2])])) or \
        np.allclose(y2[0].to_numpy(), 2.5 * X[0] + X[2], rtol=1e-15)


def test_transform_moments_many_variables(mcm):
    """sigmapoints.jl:63-72 at a size that runs the tile kernels: mean and covariance hit the targets."""
    d, n = 96, 20000
    rng = np.random.default_rng(3)
    Aq = rng.normal(size=(d, d))
    S = Aq @ Aq.T / d + np.eye(d)
    m = np.linspace(-1, 1, d)
    Y = mcm.mvnormal_particles(n, mcm.MvNormal(m, S), seed=4)
    assert len(Y) == d and Y[0].N == n
    assert np.max(np.abs(np.array(mcm.pmean(Y)) - m)) <= 1e-12
    assert np.max(np.abs(mcm.pcov(Y) - S)) <= 1e-10


# ---- shared-index gather over many variables -------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_bootstrap_shared_indices_many_variables(mcm, orc, dt):
    M, n, nd = 11, 5003, 7001
    rng = np.random.default_rng(12)
    X = rng.normal(size=(M, n)).astype(dt)
    v = [mcm.Particles(r) for r in X]
    idx = orc.rand_indices(nd, n, seed=5)
    out = mcm.bootstrap(v, nd, seed=5)
    for j in range(M):
        assert bits_equal(out[j].to_numpy(), X[j][idx]), j


# ---- batched quantiles -------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 777, 2048, 2049, 100003, 300007])
def test_quantile_batch_bit_exact(mcm, orc, dt, n):
    M = 7
    rng = np.random.default_rng(n)
    X = rng.normal(1.0, 2.0, size=(M, n)).astype(dt)
    X[1] = np.round(X[1])                   # heavy ties: thousands of identical keys around the ranks
    X[2] = 3.25                             # constant
    X[3] = np.abs(X[3]) * 1e-300 if dt == np.float64 else np.abs(X[3]) * 1e-30   # tiny magnitudes (exponent bits)
    X[4, : n // 2] = -0.0; X[4, n // 2:] = 0.0
    v = [mcm.Particles(r) for r in X]
    q = [0.025, 0.5, 0.975]
    got = mcm.pquantile(v, q)
    for j in range(M):
        assert bits_equal(got[j], orc.quantile(X[j], q)), (j, got[j], orc.quantile(X[j], q))
    q5 = [0.0, 0.1, 0.5, 0.9, 1.0]          # two sweeps of <= 4 probabilities
    got5 = mcm.pquantile(v, q5)
    for j in range(M):
        assert bits_equal(got5[j], orc.quantile(X[j], q5)), j
        assert bits_equal(got5[j], mcm.pquantile(v[j], q5))          # single-variable path agrees
    assert bits_equal(mcm.pquantile(v, 0.5), got5[:, 2])


def test_quantile_batch_bracketing_is_exact_on_hostile_layouts(mcm, orc):
    """n >= 131072 runs the sampled-bracket stage; whatever the layout of the data the result is the exact order
    statistic (a rank that escapes its bracket is re-done by the general path)."""
    n = 262147
    rng = np.random.default_rng(2)
    base = rng.normal(size=n)
    t = np.arange(n)
    X = np.stack([
        np.sort(base),                                   # sorted ascending (the unpermuted systematic sample)
        np.sort(base)[::-1].copy(),                      # sorted descending
        np.where((t // 128) % 2 == 0, base, base + 50),  # period matched to the sampling runs
        np.where(t < n // 2, 0.0, base),                 # half the keys identical
        rng.standard_cauchy(n),                          # heavy tails: huge dynamic range of keys
        1e6 + 1e-6 * base,                               # 40 leading key bits shared
        np.exp(20 * base),                               # exponent bits spread widely
    ])
    v = [mcm.Particles(r) for r in X]
    for q in ([0.025, 0.5, 0.975], [0.0, 1.0], [0.5, 0.5000001, 0.4999999, 0.25], [1e-6, 1 - 1e-6]):
        got = mcm.pquantile(v, q)
        for j in range(len(v)):
            assert bits_equal(got[j], orc.quantile(X[j], q)), (j, q, got[j], orc.quantile(X[j], q))
    # float32 and a single long variable
    x32 = rng.normal(size=1_000_003).astype(np.float32)
    assert bits_equal(mcm.pquantile([mcm.Particles(x32)], [0.1, 0.5, 0.9])[0], orc.quantile(x32, [0.1, 0.5, 0.9]))
    assert mcm._lib.lib().mcm_quantile_batch_fallbacks() >= 0


def test_quantile_batch_errors(mcm):
    v = [mcm.Particles(np.array([1.0, np.nan, 3.0] * 1000)), mcm.Particles(np.arange(3000.0))]
    with pytest.raises(mcm.MCMError):
        mcm.pquantile(v, 0.5)               # Statistics.quantile throws on NaN
    with pytest.raises(mcm.MCMError):
        mcm.pquantile(v[1:], 1.5)
    with pytest.raises(mcm.ArgumentError):
        mcm.pquantile([mcm.Particles(np.zeros(3)), mcm.Particles(np.zeros(4))], 0.5)


# ---- M x M dense steps (dense.cu) ------------------------------------------------------------------------
@pytest.fixture(params=[2, 1, 0], ids=["dense-v3", "dense-v2", "dense-v1"])
def dense_mode(request, mcm):
    """All generations of the dense kernels (mcm_dense_use_v2): the default (panelled diagonal block), the second and
    the first versions."""
    mcm._lib.lib().mcm_dense_use_v2(request.param)
    yield request.param
    mcm._lib.lib().mcm_dense_use_v2(2)


@pytest.mark.parametrize("M", [1, 5, 64, 65, 200, 517])
def test_device_cholesky_and_triangular_solve(mcm, M, dense_mode):
    """cholesky(C).L and B / L on device against LAPACK (numpy): same factor up to rounding."""
    rng = np.random.default_rng(M)
    Q = rng.normal(size=(M, M + 3))
    A = Q @ Q.T / M + 0.5 * np.eye(M)
    Ld = mcm.cholesky_lower(A)
    Lh = np.linalg.cholesky(A)
    scale = np.abs(Lh).max()
    assert np.all(np.triu(Ld, 1) == 0)
    assert np.max(np.abs(Ld - Lh)) <= 1e-13 * M * scale
    assert np.max(np.abs(Ld @ Ld.T - A)) <= 1e-14 * M * np.abs(A).max()
    B = np.tril(rng.normal(size=(M, M)))
    X = mcm.solve_right_lower(Lh, B)
    assert np.max(np.abs(X @ Lh - B)) <= 1e-12 * M * max(1.0, np.abs(X).max())
    Xr = mcm.solve_right_lower(Lh, rng.normal(size=(3, M)))                      # general right-hand sides
    assert Xr.shape == (3, M)
    with pytest.raises(mcm.MCMError):
        Bad = A.copy(); Bad[M // 2, M // 2] = -1.0
        mcm.cholesky_lower(Bad)                                                   # not positive definite -> error, like LAPACK


def test_transform_moments_device_matches_host_linear_algebra(mcm):
    """mcm_transform_moments (all on device) against the same algebra done with LAPACK on the host."""
    d, n = 70, 9000
    rng = np.random.default_rng(21)
    cols = mcm.systematic_batch(n, d, seed=5)
    Aq = rng.normal(size=(d, d)); S = Aq @ Aq.T / d + np.eye(d); m = rng.normal(size=d)
    Y = mcm.transform_moments(cols, m, S)
    X = np.stack([c.to_numpy() for c in cols])
    mu = X.mean(axis=1)
    xl = np.linalg.cholesky(np.cov(X))
    A = np.linalg.solve(xl.T, np.linalg.cholesky(S).T).T
    ref = m[:, None] + A @ (X - mu[:, None])
    Yh = np.stack([y.to_numpy() for y in Y])
    assert np.max(np.abs(Yh - ref)) <= 1e-11 * np.abs(ref).max()
    assert np.max(np.abs(np.cov(Yh) - S)) <= 1e-11 and np.max(np.abs(Yh.mean(axis=1) - m)) <= 1e-12
    Yl = mcm.transform_moments(cols, m, S, preserve_latin=True)                   # xl = Diagonal(std): sigmapoints.jl:66-67
    Al = np.linalg.cholesky(S) / X.std(axis=1, ddof=1)[None, :]
    refl = m[:, None] + Al @ (X - mu[:, None])
    assert np.max(np.abs(np.stack([y.to_numpy() for y in Yl]) - refl)) <= 1e-11 * np.abs(refl).max()
    with pytest.raises(mcm.MCMError):
        mcm.transform_moments(cols, m, -S)
    with pytest.raises(mcm.ArgumentError):
        mcm.transform_moments(cols, m, S[:5, :5])


def test_transform_moments_tensor_path_with_shared_statistics(mcm):
    """d = 256 variables: the Gram runs on tcgen05 and hands its row statistics to the tcgen05 application (no third
    pass over the particles).  Y must have exactly the requested moments
    and agree with the same transform assembled from separate calls (pcov -> host factorisations -> pgemm)."""
    d, n = 256, 20_000
    rng = np.random.default_rng(3)
    cols = mcm.systematic_batch(n, d, permute=True, seed=5)
    Aq = rng.normal(size=(d, d))
    S = Aq @ Aq.T / d + np.eye(d)
    m = np.linspace(-2, 2, d)
    Y = mcm.transform_moments(cols, m, S)
    L = mcm._lib
    assert L.lib().mcm_gram_last_engine() == 1 and L.lib().mcm_pgemm_last_engine() == 1
    assert np.max(np.abs(np.array(mcm.pmean(Y)) - m)) < 1e-13
    assert np.max(np.abs(mcm.pcov(Y) - S)) < 1e-11
    # the same transform from separate calls
    cov = mcm.pcov(cols)
    mu = np.array(mcm.pmean(cols))
    xl = np.linalg.cholesky(cov)
    A = np.tril(np.linalg.solve(xl.T, np.linalg.cholesky(S).T).T)
    Y2 = mcm.pgemm(A, cols, shift=mu, bias=m)
    for j in (0, 1, d // 2, d - 1):
        a, b = Y[j].to_numpy(), Y2[j].to_numpy()
        assert np.max(np.abs(a - b)) < 1e-10 * max(1.0, np.max(np.abs(b)))
    # twice in a row: the cached statistics and the pool are reusable
    Y3 = mcm.transform_moments(cols, m, S)
    assert np.array_equal(Y3[7].to_numpy(), Y[7].to_numpy())
    # (runs with the default, second-generation dense steps; scripts/check_dense_v2.py compares the two generations on
    # this very problem: max difference 8.9e-16)


def test_panelled_cholesky_block_is_bit_identical_to_the_column_version(mcm):
    """k_potrf_diag_v3 applies the same updates in the same order as the per-column kernel: same bits, any size,
    and the same 1-based pivot report for a matrix that is not positive definite."""
    lib = mcm._lib.lib()
    rng = np.random.default_rng(3)
    try:
        for M in (1, 7, 16, 17, 40, 64, 65, 130, 517, 1024):
            Q = rng.normal(size=(M, M + 3))
            A = Q @ Q.T + 0.1 * np.eye(M)
            lib.mcm_dense_use_v2(2); L3 = mcm.cholesky_lower(A)
            lib.mcm_dense_use_v2(1); L2 = mcm.cholesky_lower(A)
            assert np.array_equal(L3, L2), M
            assert np.max(np.abs(L3 @ L3.T - A)) <= 1e-12 * np.abs(A).max()
        B = np.eye(70); B[50, 50] = -1.0
        for mode in (2, 1):
            lib.mcm_dense_use_v2(mode)
            with pytest.raises(mcm.ArgumentError) as ei:
                mcm.cholesky_lower(B)
            assert "51" in str(ei.value)
    finally:
        lib.mcm_dense_use_v2(2)
