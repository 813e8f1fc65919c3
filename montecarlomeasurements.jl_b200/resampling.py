"""bootstrap and systematic resampling (src/resampling.jl:45-88, src/optimize.jl:22-23)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import particles as P


def rand_indices(n: int, N: int, seed: int = 0, offset: int = 0) -> P.DeviceIndex:
    """[rand(rng, 1:N) for _ in 1:n] (resampling.jl:77,81) as 0-based device indices from a
    counter-based generator (Philox-4x32-10; the stream differs from Julia's RNG, H5)."""
    d = P.DeviceIndex(n)
    L.check(L.lib().mcm_rand_indices(C.c_void_p(d.ptr), int(n), int(N), int(seed), int(offset)))
    return d


def _as_index(inds, n_expected=None) -> P.DeviceIndex:
    if isinstance(inds, P.DeviceIndex):
        return inds
    return P._upload_i64(np.asarray(inds, dtype=np.int64))


def bootstrap(p, n: int | None = None, *, seed: int = 0, inds=None):
    """bootstrap([rng,] p, n = nparticles(p)) (resampling.jl:76-88).

    p: Particles, or a list/array of Particles — then ONE index vector is drawn and applied to every
    element (resampling.jl:80-84).  `inds` (0-based) supplies the indices explicitly (e.g. from Julia's rng).
    """
    single = isinstance(p, P.Particles)
    ps = [p] if single else P._check_mv(P.flat_particles(p))
    N, dt = ps[0].N, ps[0].dtype
    n = N if n is None else int(n)
    idx = _as_index(inds) if inds is not None else rand_indices(n, N, seed)
    if idx.n != n:
        raise L.ArgumentError(L.MCM_ERR_INVALID, "index vector length must equal n")
    outs = P._empty_many(len(ps), n, dt)
    if len(ps) == 1:
        L.check(L.lib().mcm_gather(dt, C.c_void_p(ps[0].ptr), C.c_void_p(idx.ptr), C.c_void_p(outs[0].ptr), n, N))
    else:
        M = len(ps)
        srcs = (C.c_void_p * M)(*[C.c_void_p(q.ptr) for q in ps])
        dsts = (C.c_void_p * M)(*[C.c_void_p(q.ptr) for q in outs])
        L.check(L.lib().mcm_gather_multi(dt, srcs, dsts, M, C.c_void_p(idx.ptr), n, N))
    if inds is not None:
        P.check_bounds()          # caller-supplied indices: BoundsError like p.particles[inds] (resampling.jl:77,82)
    if single:
        return outs[0]
    if isinstance(p, np.ndarray):
        out = np.empty(p.shape, dtype=object)
        for j, q in enumerate(outs):
            out.flat[j] = q
        return out
    return outs


def resample_indices(weights, u: float) -> P.DeviceIndex:
    """Systematic resampling indices for unnormalised weights (restates resampling.jl:45-67 with an exact
    integer prefix scan; see include/mcm_b200.h).  weights: Float64 Particles or array; u in [0,1)."""
    w = weights if isinstance(weights, P.Particles) else P.Particles(np.asarray(weights, dtype=np.float64))
    w = w.astype(np.float64)
    d = P.DeviceIndex(w.N)
    L.check(L.lib().mcm_resample_systematic(C.c_void_p(w.ptr), w.N, float(u), C.c_void_p(d.ptr)))
    return d


def resample(p, weights, u: float = 0.5):
    """Weighted systematic resample of p (Particles or list of them) — the `resample` of BASELINE.json's
    north_star; in the reference this exists only as the dead code of resampling.jl:29-67 and as the
    weighted `sample(...; ordered=true)` + gather of optimize.jl:22-23."""
    idx = resample_indices(weights, u)
    return bootstrap(p, idx.n, inds=idx)
