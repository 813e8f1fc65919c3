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
    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = mu, sigma

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


def is_distribution(x) -> bool:
    return isinstance(x, (Normal, Uniform, MvNormal))


# ---- systematic_sample -------------------------------------------------------------------------
def systematic_sample(N: int, d=None, *, permute=True, seed: int | None = None, perm=None,
                      i0: int = 0, n_total: int | None = None, as_particles: bool = False, n_local=None):
    """systematic_sample([rng,] N, d=Normal(0,1); permute=true) (sampling.jl:6-17).

    Returns the host vector like the reference unless `as_particles=True` (then the samples stay in HBM).
    permute=True draws a device permutation from `seed` (default 0); `perm` (0-based int64 array or
    DeviceIndex) supplies an explicit permutation instead — the way to reproduce a Julia `randperm`.
    i0 / n_total / n_local: generate only grid points [i0, i0+n_local) of an n_total-point grid (sharding).
    """
    d = Normal() if d is None else d
    N = int(N)
    n_total = N if n_total is None else int(n_total)
    n_loc = N if n_local is None else int(n_local)
    code = P._dtype_code(d.eltype)
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
    else:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED,
                            f"systematic_sample on device supports Normal and Uniform, got {type(d).__name__}")
    return out if as_particles else out.to_numpy()


def particles_from_distribution(N=None, d=None, *, permute=True, systematic=True, seed=None, perm=None):
    """Particles([rng,] N, d; permute, systematic) (types.jl:67-82) and the multivariate form (types.jl:92-100)."""
    if P_is_dist(N):
        N, d = None, N
    N = P.DEFAULT_NUM_PARTICLES if N is None else int(N)
    d = Normal() if d is None else d
    if isinstance(d, MvNormal):
        return mvnormal_particles(N, d, seed=seed)
    if not systematic:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "iid sampling (systematic=false) uses the host RNG: "
                            "draw with numpy and pass the vector to Particles(v)")
    return systematic_sample(N, d, permute=permute, seed=seed, perm=perm, as_particles=True)


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
    outs = [P._empty_like(X[0]) for _ in range(M)]
    pin = (C.c_void_p * M)(*[C.c_void_p(x.ptr) for x in X])
    pout = (C.c_void_p * M)(*[C.c_void_p(y.ptr) for y in outs])
    L.check(L.lib().mcm_transform_moments(X[0].dtype, pin, M, X[0].N, m.ctypes.data_as(L._dp), S.ctypes.data_as(L._dp),
                                          1 if preserve_latin else 0, pout))
    return outs


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
    outs = [P._empty(n_loc, code) for _ in range(M)]
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
