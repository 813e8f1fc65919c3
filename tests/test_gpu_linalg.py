"""mcm_batched_linalg (per-particle small-matrix operations) on device:
  * bit-equal to the host twin of the same source (tests/test_small_linalg_twin.py pins the twin against LAPACK),
  * the reference's own assertions for exp / eigvals / lyap on particle matrices (test/runtests.jl:490-523),
  * a full-size run (N = 1e6 particles) checked through properties (A * inv(A) = I, exp(A) exp(-A) = I, trace)."""
import numpy as np
import pytest

from test_small_linalg_twin import build_twin, twin_run

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def twin():
    return build_twin()


def table(mcm, X):
    """X: (N, r, c) samples -> r x c nested list of Particles."""
    return [[mcm.Particles(np.ascontiguousarray(X[:, i, j])) for j in range(X.shape[2])] for i in range(X.shape[1])]


def gather(grid):
    return np.stack([np.stack([p.to_numpy() for p in row], axis=1) for row in grid], axis=1)


@pytest.mark.parametrize("n", list(range(1, 9)) + [9, 12, 16])
def test_device_equals_twin(mcm, twin, n):
    la = mcm.linalg
    rng = np.random.default_rng(100 + n)
    N = 3000 if n <= 8 else 1500
    A = rng.normal(size=(N, n, n))
    B = rng.normal(size=(N, n, 2))
    At, Bt = table(mcm, A), table(mcm, B)
    assert np.array_equal(gather(la.solve(At, Bt)), twin_run(twin, "solve", A, B, 2, N * n * 2).reshape(N, n, 2))
    assert np.array_equal(gather(la.inv(At)), twin_run(twin, "inv", A, nout=N * n * n).reshape(N, n, n))
    assert np.array_equal(la.det(At).to_numpy(), twin_run(twin, "det", A, nout=N))
    As = A * rng.choice([0.01, 0.3, 1.0, 3.0, 20.0], size=(N, 1, 1))
    assert np.array_equal(gather(la.expm(table(mcm, As))), twin_run(twin, "exp", As, nout=N * n * n).reshape(N, n, n))
    re, im = la.eigvals(At)
    wr, wi = twin_run(twin, "eigvals", A, nout=N * n, im=True)
    assert np.array_equal(np.stack([p.to_numpy() for p in re], 1), wr.reshape(N, n))
    assert np.array_equal(np.stack([p.to_numpy() for p in im], 1), wi.reshape(N, n))
    S = A + A.transpose(0, 2, 1)
    ws = np.stack([p.to_numpy() for p in la.eigvals_sym(table(mcm, S))], 1)
    assert np.array_equal(ws, twin_run(twin, "eigvals_sym", S, nout=N * n).reshape(N, n))
    sv = np.stack([p.to_numpy() for p in la.svdvals(At)], 1)
    assert np.array_equal(sv, twin_run(twin, "svdvals", A, nout=N * n).reshape(N, n))
    if n <= 8:
        Ast = A - max(3.0, 1.5 * np.sqrt(n) + 1.5) * np.eye(n)
        Cm = rng.normal(size=(N, n, n))
        X = gather(la.lyap(table(mcm, Ast), table(mcm, Cm)))
        assert np.array_equal(X, twin_run(twin, "lyap", Ast, Cm, n, N * n * n).reshape(N, n, n))
    mcm.check_domain()            # no singular matrix, no failed iteration among the random batch


def test_reference_assertions(mcm):
    """test/runtests.jl:490-523 restated: B = A .± 0 behaves like A; eigvals of the rotation generator; lyap(I, I)."""
    import scipy.linalg as sl
    la = mcm.linalg
    rng = np.random.default_rng(7)
    A = rng.normal(size=(2, 2))
    Bm = [[mcm.pm(A[i, j], 0.0) for j in range(2)] for i in range(2)]
    Em = la.expm(Bm)
    ref = sl.expm(A)
    assert sum(abs(mcm.pmean(Em[i][j]) - ref[i, j]) for i in range(2) for j in range(2)) < 1e-9
    re, im = la.eigvals(Bm)
    w = np.sort_complex(np.linalg.eigvals(A))
    got = np.sort_complex(np.array([mcm.pmean(r) + 1j * mcm.pmean(i_) for r, i_ in zip(re, im)]))
    assert np.sum(np.abs(np.abs(got) - np.abs(w))) < 1e-9
    # eigvals([0 1 ± 0.001; -1. 0]) sorted by imaginary part ~ [-(1 ± 0.0005)im, +(1 ± 0.0005)im]
    re, im = la.eigvals([[0.0, mcm.pm(1.0, 0.001)], [-1.0, 0.0]])
    assert abs(mcm.pmean(im[0]) + 1.0) < 0.002 and abs(mcm.pmean(im[1]) - 1.0) < 0.002
    assert abs(mcm.pstd(im[1]) - 0.0005) < 1e-5 and mcm.pmaximum(abs(re[0])) == 0.0
    # e = eigvals([1 ± 0.001 0; 0 1.]): both eigenvalues ~ 1.0 ± 0.00058 within atol 1e-2
    re, im = la.eigvals([[mcm.pm(1.0, 0.001), 0.0], [0.0, 1.0]])
    assert all(abs(mcm.pmean(r) - 1.0) < 1e-2 for r in re) and all(mcm.pmaximum(abs(i_)) == 0.0 for i_ in im)
    # lyap(pp, pp) == lyap(I, I) with pp = I .± 0.0
    pp = [[mcm.pm(1.0, 0.0), mcm.pm(0.0, 0.0)], [mcm.pm(0.0, 0.0), mcm.pm(1.0, 0.0)]]
    X = la.lyap(pp, pp)
    ref = sl.solve_continuous_lyapunov(np.eye(2), -np.eye(2))
    for i in range(2):
        for j in range(2):
            assert np.all(X[i][j].to_numpy() == ref[i, j])


def test_errors_and_flags(mcm):
    la = mcm.linalg
    L = mcm._lib
    p = mcm.pm(1.0, 0.1)
    with pytest.raises(mcm.ArgumentError):
        la.inv([[p, p, p], [p, p, p]])                       # not square
    with pytest.raises(mcm.MethodError):
        la.det([[1.0, 2.0], [3.0, 4.0]])                     # no particles
    with pytest.raises(mcm.MethodError):
        la.det([[p] * 17 for _ in range(17)])                # n > 16
    with pytest.raises(mcm.MethodError):
        la.lyap([[p] * 9 for _ in range(9)], [[p] * 9 for _ in range(9)])    # lyap handles n <= 8
    mcm.check_domain()
    la.inv([[p, p], [p, p]])                                 # exactly singular (zero pivot) for most particles
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()
    d = la.det([[p, p], [p, p]])
    assert np.all(np.abs(d.to_numpy()) < 1e-15)             # l = p * (1/p) may be 1 - ulp, as in LAPACK's getf2
    mcm.check_domain()                                       # det of a singular matrix is a value, not an error


def test_full_size_properties(mcm):
    """N = 1e6 particles, 3 x 3: 72 MB in, 72 MB out per call; checked through identities instead of a CPU loop."""
    la = mcm.linalg
    N, n = 1_000_000, 3
    A = [[mcm.systematic_sample(N, mcm.Normal(1.0 if i == j else 0.2, 0.05), seed=10 * i + j, as_particles=True)
          for j in range(n)] for i in range(n)]
    Ai = la.inv(A)
    for i in range(n):
        for j in range(n):
            acc = A[i][0] * Ai[0][j]
            for k in range(1, n):
                acc = acc + A[i][k] * Ai[k][j]
            m = mcm.moments(acc - (1.0 if i == j else 0.0))
            assert max(abs(m.min), abs(m.max)) < 1e-13
    Ep = la.expm(A)
    En = la.expm([[-A[i][j] for j in range(n)] for i in range(n)])
    for i in range(n):
        for j in range(n):
            acc = Ep[i][0] * En[0][j]
            for k in range(1, n):
                acc = acc + Ep[i][k] * En[k][j]
            m = mcm.moments(acc - (1.0 if i == j else 0.0))
            assert max(abs(m.min), abs(m.max)) < 1e-12
    # det(exp(A)) = exp(tr A);  sum of eigenvalues = trace
    lhs = la.det(Ep)
    tr = A[0][0] + A[1][1] + A[2][2]
    m = mcm.moments(lhs / mcm.exp(tr) - 1.0)
    assert max(abs(m.min), abs(m.max)) < 1e-12
    re, im = la.eigvals(A)
    m = mcm.moments((re[0] + re[1] + re[2]) - tr)
    assert max(abs(m.min), abs(m.max)) < 1e-12
    m = mcm.moments(im[0] + im[1] + im[2])
    assert max(abs(m.min), abs(m.max)) == 0.0


# ---- matrices with Complex{Particles} entries; log of a real uncertain matrix (complex.jl:104-181, particles.jl:603) -------
def ctable(mcm, Z):
    return [[mcm.ComplexParticles(mcm.Particles(np.ascontiguousarray(Z[:, i, j].real)), mcm.Particles(np.ascontiguousarray(Z[:, i, j].imag)))
             for j in range(Z.shape[2])] for i in range(Z.shape[1])]


def cgather(grid):
    return np.stack([np.stack([z.re.to_numpy() + 1j * z.im.to_numpy() for z in row], axis=1) for row in grid], axis=1)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 8, 9, 12, 16])
def test_complex_matrix_functions_equal_the_twin(mcm, n):
    """Device == host twin bit for bit (tests/test_small_linalg_c_twin.py pins the twin against LAPACK / scipy)."""
    import test_small_linalg_c_twin as tw
    twin = tw.build_twin()
    la = mcm.linalg
    rng = np.random.default_rng(300 + n)
    N = 2000
    Z = rng.normal(size=(N, n, n)) + 1j * rng.normal(size=(N, n, n))
    Zt = ctable(mcm, Z)
    assert np.array_equal(cgather(la.cexpm(ctable(mcm, 0.7 * Z))), tw.twin_run(twin, "exp", 0.7 * Z))
    d = la.cdet(Zt)
    assert np.array_equal(d.re.to_numpy() + 1j * d.im.to_numpy(), tw.twin_run(twin, "det", Z))
    w = la.ceigvals(Zt)
    assert np.array_equal(np.stack([z.re.to_numpy() + 1j * z.im.to_numpy() for z in w], 1), tw.twin_run(twin, "eigvals", Z))
    sv = la.csvdvals(Zt)
    assert np.array_equal(np.stack([p.to_numpy() for p in sv], 1), tw.twin_run(twin, "svdvals", Z))
    import scipy.linalg as sl
    B = np.array([sl.expm(0.5 * z) for z in Z[:500]])
    assert np.array_equal(cgather(la.logm(ctable(mcm, B))), tw.twin_run(twin, "log", B))
    R = rng.normal(size=(500, n, n))                                   # real input: A_im == NULL
    Lr = cgather(la.logm(table(mcm, R)))
    assert np.array_equal(Lr, tw.twin_run(twin, "log", R))
    mcm.check_domain()


def test_complex_matrix_reference_style_assertions(mcm):
    """B = A .± 0 behaves like A (the pattern of runtests.jl:490-523 and :572-575 for complex entries); exp(log(A)) == A;
    log of a positive definite uncertain matrix is real; a singular matrix raises the DomainError flag."""
    import scipy.linalg as sl
    la = mcm.linalg
    rng = np.random.default_rng(5)
    A = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    N = 64
    Zt = [[mcm.ComplexParticles(mcm.Particles(np.full(N, A[i, j].real)), mcm.Particles(np.full(N, A[i, j].imag))) for j in range(3)] for i in range(3)]
    E = cgather(la.cexpm(Zt))
    assert np.max(np.abs(E - sl.expm(A)[None])) < 1e-12 * np.abs(sl.expm(A)).max()
    Lg = cgather(la.logm(Zt))
    assert np.max(np.abs(Lg - sl.logm(A)[None])) < 1e-11 * np.linalg.cond(A)
    assert abs(complex(la.cdet(Zt).re.to_numpy()[0], la.cdet(Zt).im.to_numpy()[0]) - np.linalg.det(A)) < 1e-12 * abs(np.linalg.det(A))
    sv = np.array([p.to_numpy()[0] for p in la.csvdvals(Zt)])
    assert np.max(np.abs(sv - np.linalg.svd(A, compute_uv=False))) < 1e-13 * sv[0]
    w = np.array([complex(z.re.to_numpy()[0], z.im.to_numpy()[0]) for z in la.ceigvals(Zt)])
    wr = np.linalg.eigvals(A)
    assert np.max(np.min(np.abs(w[:, None] - wr[None, :]), axis=1)) < 1e-12 * np.abs(wr).max()
    # mixed entries: plain numbers, Particles and ComplexParticles promote like the reference's `promote`
    p = mcm.pm(1.0, 0.1, seed=1)
    M = [[p, 2.0 + 1.0j], [0.5, mcm.ComplexParticles(p, p)]]
    ph = p.to_numpy()
    Zs = np.zeros((p.N, 2, 2), dtype=complex)
    Zs[:, 0, 0] = ph; Zs[:, 0, 1] = 2 + 1j; Zs[:, 1, 0] = 0.5; Zs[:, 1, 1] = ph + 1j * ph
    dd = la.cdet(M)
    assert np.max(np.abs(dd.re.to_numpy() + 1j * dd.im.to_numpy() - np.linalg.det(Zs))) < 1e-13
    # exp(log(A)) == A on uncertain real matrices, and real logs for positive definite ones
    R = rng.normal(size=(4000, 3, 3))
    S = R @ R.transpose(0, 2, 1) + 0.5 * np.eye(3)
    Ls = cgather(la.logm(table(mcm, S)))
    assert np.max(np.abs(Ls.imag)) < 1e-12
    back = cgather(la.cexpm(ctable(mcm, Ls)))
    assert np.max(np.abs(back - S) / np.abs(S).max(axis=(1, 2), keepdims=True)) < 1e-11
    mcm.check_domain()
    la.logm([[mcm.Particles(np.zeros(8)), 0.0], [0.0, mcm.Particles(np.ones(8))]])
    with pytest.raises(mcm.DomainError):
        mcm.check_domain()
    with pytest.raises(mcm.MethodError):
        la.logm(np.zeros((2, 2)))


@pytest.mark.parametrize("n", [3, 8, 12])
def test_float32_particles_round_the_float64_result(mcm, twin, n):
    """Float32 entries: the per-thread algorithm runs in Float64 on the widened inputs and the result is rounded once
    (the reference's host loop would run LAPACK in Float32; documented deviation, tighter)."""
    la = mcm.linalg
    rng = np.random.default_rng(700 + n)
    N = 1000
    A = (rng.normal(size=(N, n, n)) + 3.0 * np.eye(n)).astype(np.float32)
    B = rng.normal(size=(N, n, 1)).astype(np.float32)
    At, Bt = table(mcm, A), table(mcm, B)
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    got = gather(la.solve(At, Bt))
    assert got.dtype == np.float32
    assert np.array_equal(got, twin_run(twin, "solve", A64, B64, 1, N * n).reshape(N, n, 1).astype(np.float32))
    assert np.array_equal(la.det(At).to_numpy(), twin_run(twin, "det", A64, nout=N).astype(np.float32))
    assert np.array_equal(gather(la.expm(table(mcm, 0.2 * A))), twin_run(twin, "exp", (np.float32(0.2) * A).astype(np.float64), nout=N * n * n).reshape(N, n, n).astype(np.float32))
    sv = np.stack([p.to_numpy() for p in la.svdvals(At)], 1)
    assert np.array_equal(sv, twin_run(twin, "svdvals", A64, nout=N * n).reshape(N, n).astype(np.float32))
    mcm.check_domain()
