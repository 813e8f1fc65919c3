"""systematic_sample and the sampling constructors (src/sampling.jl:6-17,
src/types.jl:67-100, src/particles.jl:20-50, src/sigmapoints.jl:63-72).

The quantile grid and the inverse normal CDF are evaluated on device
(mcm_systematic_normal / mcm_systematic_uniform); the random permutation of
sampling.jl:11 is either supplied by the caller (bit-exact by construction, H5)
or drawn on device from a counter-based generator (mcm_randperm).
"""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from . import _lib as L
from . import particles as P


# ---- minimal Distributions.jl stand-ins (only what the hot path constructs) -----------------
class Normal:
    def __init__(self, mu=0.0, sigma=None):
        if sigma is None:                      # Normal(μ) = Normal(μ, one(μ)): Normal(0.1f0) stays Float32
            sigma = np.float32(1.0) if isinstance(mu, np.float32) else 1.0
        self.mu, self.sigma = mu, sigma

    def __repr__(self):
        return f"Normal({self.mu!r}, {self.sigma!r})"

    @property
    def eltype(self):
        # T = eltype(mean(d)) (sampling.jl:7); Normal(μ,σ) promotes its parameters to float
        both32 = all(isinstance(x, np.float32) for x in (self.mu, self.sigma))
        return np.float32 if both32 else np.float64


class Uniform:
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = a, b

    @property
    def eltype(self):
        both32 = all(isinstance(x, np.float32) for x in (self.a, self.b))
        return np.float32 if both32 else np.float64


class MvNormal:
    """MvNormal(μ, Σ); a scalar σ means σ²·I (std), a vector means the variances — as Distributions does
    for `μ ± σ` with vector μ (particles.jl:5-6,21)."""

    def __init__(self, mu, S):
        self.mu = np.asarray(mu, dtype=np.float64)
        d = len(self.mu)
        if isinstance(S, numbers.Real):
            self.S = np.eye(d) * float(S) ** 2
        else:
            S = np.asarray(S, dtype=np.float64)
            self.S = np.diag(S) if S.ndim == 1 else S


# ---- the other continuous univariate families (Distributions.jl parameterisation; csrc/mcm_dist.cuh) ----------
class UnivariateDistribution:
    """A continuous univariate Distribution whose quantile function runs on device (mcm_systematic_dist)."""
    kind = None          # MCM_D_* name

    def __init__(self, *params):
        self._params = tuple(params)

    @property
    def params(self):
        p = [float(x) for x in self._params] + [0.0, 0.0, 0.0]
        return p[0], p[1], p[2]

    @property
    def eltype(self):
        return np.float32 if self._params and all(isinstance(x, np.float32) for x in self._params) else np.float64

    def __repr__(self):
        return f"{type(self).__name__}{self._params}"


def _family(name, kind, defaults, check=None):
    def __init__(self, *args):
        vals = list(args) + list(defaults[len(args):])
        if len(args) > len(defaults):
            raise L.MethodError(L.MCM_ERR_UNSUPPORTED, f"{name} takes at most {len(defaults)} parameters")
        if check is not None and not check(*[float(v) for v in vals]):
            raise L.ArgumentError(L.MCM_ERR_INVALID, f"{name}{tuple(vals)}: parameters outside the distribution's domain")
        UnivariateDistribution.__init__(self, *vals)
    return type(name, (UnivariateDistribution,), {"kind": kind, "__init__": __init__,
                                                  "__doc__": f"{name}{tuple(defaults)} as in Distributions.jl; quantile on device (MCM_D_{kind})."})


Exponential = _family("Exponential", "EXPONENTIAL", (1.0,), lambda t: t > 0)
LogNormal = _family("LogNormal", "LOGNORMAL", (0.0, 1.0), lambda m, s: s >= 0)
Cauchy = _family("Cauchy", "CAUCHY", (0.0, 1.0), lambda m, s: s > 0)
Laplace = _family("Laplace", "LAPLACE", (0.0, 1.0), lambda m, t: t > 0)
Logistic = _family("Logistic", "LOGISTIC", (0.0, 1.0), lambda m, t: t > 0)
Rayleigh = _family("Rayleigh", "RAYLEIGH", (1.0,), lambda s: s > 0)
Weibull = _family("Weibull", "WEIBULL", (1.0, 1.0), lambda a, t: a > 0 and t > 0)
Pareto = _family("Pareto", "PARETO", (1.0, 1.0), lambda a, t: a > 0 and t > 0)
Gumbel = _family("Gumbel", "GUMBEL", (0.0, 1.0), lambda m, t: t > 0)
TriangularDist = _family("TriangularDist", "TRIANGULAR", (0.0, 1.0, 0.5), lambda a, b, c: a <= c <= b and a < b)
Frechet = _family("Frechet", "FRECHET", (1.0, 1.0), lambda a, t: a > 0 and t > 0)
Kumaraswamy = _family("Kumaraswamy", "KUMARASWAMY", (1.0, 1.0), lambda a, b: a > 0 and b > 0)
Gamma = _family("Gamma", "GAMMA", (1.0, 1.0), lambda a, t: a > 0 and t > 0)
Erlang = _family("Erlang", "GAMMA", (1.0, 1.0), lambda a, t: a > 0 and t > 0 and float(a).is_integer())
Chisq = _family("Chisq", "CHISQ", (1.0,), lambda n: n > 0)
InverseGamma = _family("InverseGamma", "INVERSEGAMMA", (1.0, 1.0), lambda a, t: a > 0 and t > 0)
Beta = _family("Beta", "BETA", (1.0, 1.0), lambda a, b: a > 0 and b > 0)
TDist = _family("TDist", "TDIST", (1.0,), lambda n: n > 0)
FDist = _family("FDist", "FDIST", (1.0, 1.0), lambda a, b: a > 0 and b > 0)


class GenericDistribution:
    """Any other univariate distribution, through a HOST-SUPPLIED QUANTILE TABLE: `quantile` maps a numpy vector of
    probabilities to values (a scipy.stats frozen distribution's `.ppf` works as is).  The systematic grid is
    evaluated on the host once, uploaded, and permuted on device.  `rvs(n, rng)` (optional) draws iid samples for
    systematic=False / discrete distributions (types.jl:67-82 with VS != Continuous)."""

    def __init__(self, quantile=None, rvs=None, continuous: bool = True, eltype=np.float64):
        if quantile is not None and hasattr(quantile, "ppf"):          # a scipy.stats frozen distribution
            frozen = quantile
            quantile = frozen.ppf
            rvs = rvs or (lambda n, rng: frozen.rvs(size=n, random_state=rng))
            continuous = hasattr(frozen.dist, "pdf")
        self.quantile, self.rvs, self.continuous, self.eltype = quantile, rvs, continuous, eltype


def is_distribution(x) -> bool:
    return isinstance(x, (Normal, Uniform, MvNormal, UnivariateDistribution, GenericDistribution)) or \
        (hasattr(x, "ppf") and hasattr(x, "rvs"))


def _as_distribution(d):
    if hasattr(d, "ppf") and hasattr(d, "rvs") and not isinstance(d, GenericDistribution):
        return GenericDistribution(d)
    return d


# ---- systematic_sample -------------------------------------------------------------------------
def systematic_sample(N: int, d=None, *, permute=True, seed: int | None = None, perm=None,
                      i0: int = 0, n_total: int | None = None, as_particles: bool = False, n_local=None):
    """systematic_sample([rng,] N, d=Normal(0,1); permute=true) (sampling.jl:6-17).

    Returns the host vector like the reference unless `as_particles=True` (then the samples stay in HBM).
    permute=True draws a device permutation from `seed` (default 0); `perm` (0-based int64 array or
    DeviceIndex) supplies an explicit permutation instead — the way to reproduce a Julia `randperm`.
    i0 / n_total / n_local: generate only grid points [i0, i0+n_local) of an n_total-point grid (sharding).
    """
    d = _as_distribution(Normal() if d is None else d)
    N = int(N)
    n_total = N if n_total is None else int(n_total)
    n_loc = N if n_local is None else int(n_local)
    code = P._dtype_code(d.eltype)
    if isinstance(d, GenericDistribution):
        return _systematic_generic(d, N, n_total, n_loc, i0, permute, seed, perm, as_particles)
    out = P._empty(n_loc, code)
    perm_dev = None
    if perm is not None:
        perm_dev = perm if isinstance(perm, P.DeviceIndex) else P._upload_i64(np.asarray(perm))
        if perm_dev.n != n_loc:
            raise L.ArgumentError(L.MCM_ERR_INVALID, "permutation length must equal N")
    elif permute:
        perm_dev = P.DeviceIndex(n_loc)
        if n_loc != n_total:
            raise L.ArgumentError(L.MCM_ERR_INVALID, "sharded generation needs an explicit perm or permute=False")
        L.check(L.lib().mcm_randperm(C.c_void_p(perm_dev.ptr), n_loc, int(0 if seed is None else seed)))
    pp = C.c_void_p(perm_dev.ptr) if perm_dev is not None else None
    if isinstance(d, Normal):
        L.check(L.lib().mcm_systematic_normal(code, C.c_void_p(out.ptr), n_loc, float(d.mu), float(d.sigma), pp,
                                              int(i0), n_total))
    elif isinstance(d, Uniform):
        L.check(L.lib().mcm_systematic_uniform(code, C.c_void_p(out.ptr), n_loc, float(d.a), float(d.b), pp,
                                               int(i0), n_total))
    elif isinstance(d, UnivariateDistribution):
        a, b, c = d.params
        L.check(L.lib().mcm_systematic_dist(code, L.ENUMS["MCM_D_" + d.kind], a, b, c, C.c_void_p(out.ptr), n_loc, pp,
                                            int(i0), n_total))
    else:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, f"systematic_sample: {type(d).__name__} is not a univariate distribution")
    return out if as_particles else out.to_numpy()


def _systematic_generic(d, N, n_total, n_loc, i0, permute, seed, perm, as_particles):
    """Host-supplied quantile table: grid values from d.quantile on the host, permutation on device."""
    if d.quantile is None:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "GenericDistribution without a quantile function cannot be sampled systematically")
    i = np.arange(int(i0), int(i0) + n_loc, dtype=np.float64)
    vals = np.asarray(d.quantile((2.0 * i + 1.0) / (2.0 * n_total)), dtype=d.eltype)
    out = P.Particles(np.ascontiguousarray(vals))
    if perm is not None or permute:
        if perm is None:
            if n_loc != n_total:
                raise L.ArgumentError(L.MCM_ERR_INVALID, "sharded generation needs an explicit perm or permute=False")
            pd = P.DeviceIndex(n_loc)
            L.check(L.lib().mcm_randperm(C.c_void_p(pd.ptr), n_loc, int(0 if seed is None else seed)))
        else:
            pd = perm if isinstance(perm, P.DeviceIndex) else P._upload_i64(np.asarray(perm))
        res = P._empty(n_loc, out.dtype)
        L.check(L.lib().mcm_gather(out.dtype, C.c_void_p(out.ptr), C.c_void_p(pd.ptr), C.c_void_p(res.ptr), n_loc, n_loc))
        P.check_bounds()
        out = res
    return out if as_particles else out.to_numpy()


def rand_particles_dist(N: int, d, seed: int = 0, offset: int = 0):
    """iid draws Particles(N, d; systematic=false) (types.jl:71,79): inversion of device Philox uniforms for the
    device families; `rvs` on the host for a GenericDistribution (Julia's RNG stream is not reproducible, H5)."""
    d = _as_distribution(d)
    N = int(N)
    if isinstance(d, GenericDistribution):
        rng = np.random.default_rng(seed)
        if d.rvs is not None:
            v = np.asarray(d.rvs(N, rng))
        elif d.quantile is not None:
            v = np.asarray(d.quantile(rng.random(N)))
        else:
            raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "GenericDistribution needs rvs or quantile")
        return P.Particles(np.ascontiguousarray(v.astype(np.float64) if v.dtype.kind in "iub" else v))
    code = P._dtype_code(d.eltype)
    out = P._empty(N, code)
    if isinstance(d, Normal):
        kind, (a, b, c) = "NORMAL", (float(d.mu), float(d.sigma), 0.0)
    elif isinstance(d, Uniform):
        kind, (a, b, c) = "UNIFORM", (float(d.a), float(d.b), 0.0)
    else:
        kind, (a, b, c) = d.kind, d.params
    L.check(L.lib().mcm_rand_dist(code, L.ENUMS["MCM_D_" + kind], a, b, c, C.c_void_p(out.ptr), N, int(seed), int(offset)))
    return out


def quantile_dist(d, p):
    """quantile.(d, p) for a Particles / array of probabilities, on device."""
    d = _as_distribution(d)
    pp = p if isinstance(p, P.Particles) else P.Particles(np.asarray(p, dtype=np.float64))
    if isinstance(d, Normal):
        kind, (a, b, c) = "NORMAL", (float(d.mu), float(d.sigma), 0.0)
    elif isinstance(d, Uniform):
        kind, (a, b, c) = "UNIFORM", (float(d.a), float(d.b), 0.0)
    elif isinstance(d, UnivariateDistribution):
        kind, (a, b, c) = d.kind, d.params
    else:
        return P.Particles(np.asarray(d.quantile(pp.to_numpy())))
    out = P._empty_like(pp)
    L.check(L.lib().mcm_quantile_dist(pp.dtype, L.ENUMS["MCM_D_" + kind], a, b, c, C.c_void_p(pp.ptr), C.c_void_p(out.ptr), pp.N))
    return out


def particles_from_distribution(N=None, d=None, *, permute=True, systematic=None, seed=None, perm=None):
    """Particles([rng,] N, d; permute, systematic) (types.jl:67-82) and the multivariate form (types.jl:92-100)."""
    if P_is_dist(N):
        N, d = None, N
    N = P.DEFAULT_NUM_PARTICLES if N is None else int(N)
    d = Normal() if d is None else d
    if isinstance(d, MvNormal):
        return mvnormal_particles(N, d, seed=seed)
    d = _as_distribution(d)
    if systematic is None:                       # systematic = (VS == Continuous) by default (types.jl:67)
        systematic = getattr(d, "continuous", True)
    if not systematic:
        return rand_particles_dist(N, d, seed=0 if seed is None else int(seed))
    return systematic_sample(N, d, permute=permute, seed=seed, perm=perm, as_particles=True)


def particles_of_type(T, N=None, d=None, **kw):
    """PT(::Type{T}, N = 2000, d = Normal(T(0), T(1)); kwargs...) (types.jl:85-88): the distribution's element type must be
    T — `Particles(Float32, 250, Gamma(0.1))` throws ArgumentError in the reference (runtests.jl:241) because Gamma(0.1)
    samples Float64."""
    T = np.float64 if T is float else T
    N = P.DEFAULT_NUM_PARTICLES if N is None else int(N)
    if d is None:
        d = Normal(T(0), T(1))
    dd = _as_distribution(d)
    et = np.dtype(getattr(dd, "eltype", np.float64))
    if et != np.dtype(T):
        raise L.ArgumentError(L.MCM_ERR_INVALID, f"Element type of the provided distribution {d!r} does not match {np.dtype(T).name}. "
                              "The element type of a distribution is the element type of a value sampled from it.")
    return particles_from_distribution(N, d, **kw)


def P_is_dist(x):
    return is_distribution(x)


def pm(mu, sigma=None, N: int = P.DEFAULT_NUM_PARTICLES, seed: int | None = None, perm=None):
    """μ ± σ (particles.jl:20-22): N=2000 systematic Normal(μ,σ) particles; unary ±σ means 0 ± σ;
    vector μ -> MvNormal (particles.jl:21)."""
    if sigma is None:
        mu, sigma = 0.0 * mu, mu
    if isinstance(mu, (list, tuple, np.ndarray)):
        return mvnormal_particles(N, MvNormal(mu, sigma), seed=seed)
    # T = promote_type(float(typeof μ), float(typeof σ)) (particles.jl:20)
    dt = np.float32 if all(isinstance(x, np.float32) for x in (mu, sigma)) else np.float64
    d = Normal(dt(mu), dt(sigma))
    return systematic_sample(N, d, permute=True, seed=seed, perm=perm, as_particles=True)


def mp(mu, sigma=None, seed: int | None = None, perm=None):
    """μ ∓ σ (particles.jl:23-25): the StaticParticles flavour, N = 100 (storage is the same on device)."""
    return pm(mu, sigma, N=P.DEFAULT_STATIC_NUM_PARTICLES, seed=seed, perm=perm)


def boxtimes(a, d, seed: int | None = None):
    """a ⊠ d = a * Particles(d) (particles.jl:42): e.g. boxtimes(2.0, Gamma(1))."""
    return a * particles_from_distribution(P.DEFAULT_NUM_PARTICLES, d, seed=seed)


def boxplus(a, d, seed: int | None = None):
    """a ⊞ d = a + Particles(d) (particles.jl:50): e.g. boxplus(1, GenericDistribution(scipy.stats.binom(3, 0.5)))."""
    return a + particles_from_distribution(P.DEFAULT_NUM_PARTICLES, d, seed=seed)


def outer_product(dists, N: int = 100_000, seed: int = 0):
    """outer_product(dists, N) / μ ⊗ σ (particles.jl:60-75): Nd = floor(N^(1/d)) permuted systematic samples per
    dimension, combined as the full Cartesian product (Iterators.product: first dimension fastest)."""
    dists = list(dists)
    dim = len(dists)
    Nd = max(int(np.floor(N ** (1.0 / dim) + 1e-9)), 1)
    cols = [systematic_sample(Nd, d, permute=True, seed=seed + k) for k, d in enumerate(dists)]
    grids = np.meshgrid(*cols, indexing="ij")
    return [P.Particles(np.ascontiguousarray(g.reshape(-1, order="F"))) for g in grids]


def otimes(mu, sigma, N: int = 100_000):
    """⊗(μ, σ) = outer_product(Normal.(μ, σ)) (particles.jl:58)."""
    return outer_product([Normal(m, s) for m, s in zip(np.atleast_1d(mu), np.atleast_1d(sigma))], N)


def dotdot(a, b, N: int = P.DEFAULT_NUM_PARTICLES, seed: int | None = None):
    """a .. b (particles.jl:33): Uniform(a,b) particles."""
    return systematic_sample(N, Uniform(a, b), permute=True, seed=seed, as_particles=True)


# ---- multivariate construction (types.jl:92-100; systematic route docs/index.md:141 + sigmapoints.jl:63-72)
def transform_moments(X, m, S, preserve_latin: bool = False):
    """transform_moments(X, m, Σ) (sigmapoints.jl:63-72) on a list of M device Particles (the N x M sample):
    centre, xl = chol(cov X).L (or Diagonal(std)), Y = m + (chol(Σ).L / xl) X' — one C-ABI call, all on device:
    Gram (mcm_gram), blocked Cholesky and triangular solve (dense.cu), application (mcm_pgemm with the centring
    fused as its shift)."""
    X = P._check_mv(X)
    M = len(X)
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(m, dtype=np.float64), (M,)))
    S = np.ascontiguousarray(np.asarray(S, dtype=np.float64))
    if S.shape != (M, M):
        raise L.ArgumentError(L.MCM_ERR_INVALID, f"DimensionMismatch: Σ must be {M} x {M}")
    outs = P._empty_many(M, X[0].N, X[0].dtype)
    pin = (C.c_void_p * M)(*[C.c_void_p(x.ptr) for x in X])
    pout = (C.c_void_p * M)(*[C.c_void_p(y.ptr) for y in outs])
    L.check(L.lib().mcm_transform_moments(X[0].dtype, pin, M, X[0].N, m.ctypes.data_as(L._dp), S.ctypes.data_as(L._dp),
                                          1 if preserve_latin else 0, pout))
    return outs


def sigmapoints(m, S=None) -> np.ndarray:
    """sigmapoints(m, Σ) / sigmapoints(d::Normal) / sigmapoints(d::MvNormal) (sigmapoints.jl:35-45): the (2n+1) x n matrix
    [X; -X; 0] .+ m' with X = cholesky(n Σ).U (factorised on device, mcm_potrf).  `Particles(sigmapoints(m, Σ))` gives the
    n uncertain values (dim 1 = particles); a scalar Σ is a VARIANCE, like in the reference."""
    if isinstance(m, Normal) and S is None:
        m, S = float(m.mu), float(m.sigma) ** 2
    elif isinstance(m, MvNormal) and S is None:
        m, S = m.mu, m.S
    mv = np.atleast_1d(np.asarray(m, dtype=np.float64))
    n = mv.size
    Sm = np.asarray(S, dtype=np.float64)
    if Sm.ndim == 0:
        Sm = np.diag(np.full(n, float(Sm)))
    if Sm.shape != (n, n):
        raise L.ArgumentError(L.MCM_ERR_INVALID, f"DimensionMismatch: Σ must be {n} x {n}")
    X = cholesky_lower(n * (Sm + Sm.T) / 2.0).T             # cholesky(Symmetric(nΣ)).U
    return np.vstack([X, -X, np.zeros((1, n))]) + mv[None, :]


def cholesky_lower(A) -> np.ndarray:
    """cholesky(A).L computed on device (mcm_potrf); raises ArgumentError if A is not positive definite."""
    A = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
    M = A.shape[0]
    d = P._alloc(A.nbytes)
    try:
        L.check(L.lib().mcm_upload(C.c_void_p(d), A.ctypes.data_as(C.c_void_p), A.nbytes))
        info = C.c_int(0)
        L.check(L.lib().mcm_potrf(C.c_void_p(d), M, C.byref(info)))
        out = np.empty_like(A)
        L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(d), out.nbytes))
    finally:
        L.lib().mcm_free(C.c_void_p(d))
    return out


def solve_right_lower(Lm, B) -> np.ndarray:
    """B / L for lower-triangular L (X L = B) on device (mcm_trsm_right_lower)."""
    Lm = np.ascontiguousarray(np.asarray(Lm, dtype=np.float64))
    B = np.ascontiguousarray(np.atleast_2d(np.asarray(B, dtype=np.float64)))
    M = Lm.shape[0]
    dl, db = P._alloc(Lm.nbytes), P._alloc(B.nbytes)
    try:
        L.check(L.lib().mcm_upload(C.c_void_p(dl), Lm.ctypes.data_as(C.c_void_p), Lm.nbytes))
        L.check(L.lib().mcm_upload(C.c_void_p(db), B.ctypes.data_as(C.c_void_p), B.nbytes))
        L.check(L.lib().mcm_trsm_right_lower(C.c_void_p(dl), C.c_void_p(db), M, B.shape[0]))
        out = np.empty_like(B)
        L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(db), out.nbytes))
    finally:
        L.lib().mcm_free(C.c_void_p(dl)); L.lib().mcm_free(C.c_void_p(db))
    return out


def index_bijection(n: int, seed: int = 0, column: int = 0, i0: int = 0, n_total: int | None = None) -> P.DeviceIndex:
    """Keyed pseudo-random permutation pi(i0 .. i0+n) of [0, n_total) evaluated on device with no sort and no
    communication (mcm_index_bijection) — the sharding-friendly stand-in for randperm (sampling.jl:11)."""
    n_total = int(n if n_total is None else n_total)
    out = P.DeviceIndex(int(n))
    L.check(L.lib().mcm_index_bijection(C.c_void_p(out.ptr), int(n), int(seed), int(column), int(i0), n_total))
    return out


def systematic_batch(N: int, dists, *, permute=True, seed: int = 0, i0: int = 0, n_total: int | None = None,
                     n_local: int | None = None, dtype=np.float64):
    """M systematic columns in one launch (mcm_systematic_batch): column v is systematic_sample(N, dists[v])
    permuted by the keyed bijection of (seed, v).  `dists`: an int M (standard normals) or a list of Normal /
    Uniform (one family per call)."""
    if isinstance(dists, numbers.Integral):
        M, kind, a, b = int(dists), 0, None, None
    else:
        dists = list(dists)
        M = len(dists)
        if all(isinstance(d, Normal) for d in dists):
            kind, a, b = 0, np.array([float(d.mu) for d in dists]), np.array([float(d.sigma) for d in dists])
        elif all(isinstance(d, Uniform) for d in dists):
            kind, a, b = 1, np.array([float(d.a) for d in dists]), np.array([float(d.b) for d in dists])
        else:
            raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "systematic_batch: all Normal or all Uniform")
    n_total = int(N if n_total is None else n_total)
    n_loc = int(N if n_local is None else n_local)
    code = P._dtype_code(dtype)
    outs = P._empty_many(M, n_loc, code)
    ptrs = (C.c_void_p * M)(*[C.c_void_p(o.ptr) for o in outs])
    L.check(L.lib().mcm_systematic_batch(code, kind, ptrs, M, n_loc,
                                         None if a is None else a.ctypes.data_as(L._dp),
                                         None if b is None else b.ctypes.data_as(L._dp),
                                         1 if permute else 0, int(seed), int(i0), n_total))
    return outs


def mvnormal_particles(N: int, d: MvNormal, seed: int | None = None):
    """Systematic multivariate Normal: independent permuted systematic columns (docs index.md:141) moved to
    (μ, Σ) by transform_moments.  (The reference default for MvNormal draws iid randn — a host RNG stream
    that cannot be reproduced here, H5; this is its documented systematic alternative.)"""
    cols = systematic_batch(N, len(d.mu), permute=True, seed=0 if seed is None else int(seed))
    return transform_moments(cols, d.mu, d.S)
