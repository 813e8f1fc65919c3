"""union / intersect (src/particles.jl:360-381) and with_nominal / nominal (src/nominal.jl), on device.

Nothing here touches the samples on the host: concatenation is two device copies, the support filter is a stable
device compaction (`mcm_filter_range`), the nearest particle an argmin reduction (`mcm_argmin_absdiff`,
`mcm_argmin_sqdist`), the replacement a two-element kernel (`mcm_shuffle_insert`)."""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from . import _lib as L
from . import particles as P


def _same_eltype(p1: P.Particles, p2: P.Particles, who: str):
    if not isinstance(p1, P.Particles) or not isinstance(p2, P.Particles):
        raise L.MethodError(L.MCM_ERR_INVALID, f"{who}: both arguments must be Particles")
    if p1.dtype != p2.dtype:
        # Base.union(p1::PT{T,NT}, p2::PT{T,NS}) is only defined for one element type T (particles.jl:366)
        raise L.MethodError(L.MCM_ERR_INVALID, f"{who}: both arguments need the same element type")


def union(p1: P.Particles, p2: P.Particles) -> P.Particles:
    """union(p1, p2) (particles.jl:366-368): `Particles([p1.particles; p2.particles])`, N = N1 + N2."""
    _same_eltype(p1, p2, "union")
    es = P._NP[p1.dtype].itemsize
    out = P._empty(p1.N + p2.N, p1.dtype)
    lib = L.lib()
    L.check(lib.mcm_copy(C.c_void_p(out.ptr), C.c_void_p(p1.ptr), p1.N * es))
    L.check(lib.mcm_copy(C.c_void_p(out.ptr + p1.N * es), C.c_void_p(p2.ptr), p2.N * es))
    return out


def intersect(p1: P.Particles, p2: P.Particles) -> P.Particles:
    """intersect(p1, p2) (particles.jl:375-380): the particles of both inside the common support
    [max(minima), min(maxima)], p1's survivors first, each in its original order.  The length is data dependent."""
    _same_eltype(p1, p2, "intersect")
    mi = max(P.pminimum(p1), P.pminimum(p2))
    ma = min(P.pmaximum(p1), P.pmaximum(p2))
    buf = P._empty(p1.N + p2.N, p1.dtype)
    lib = L.lib()
    k1, k2 = C.c_size_t(0), C.c_size_t(0)
    L.check(lib.mcm_filter_range(p1.dtype, C.c_void_p(p1.ptr), p1.N, mi, ma, C.c_void_p(buf.ptr), 0, C.byref(k1)))
    L.check(lib.mcm_filter_range(p2.dtype, C.c_void_p(p2.ptr), p2.N, mi, ma, C.c_void_p(buf.ptr), k1.value, C.byref(k2)))
    n = k1.value + k2.value
    buf.N = n                       # the allocation stays as large as it was; only the logical length shrinks
    # disjoint supports give Particles of length 0, as in the reference (runtests.jl:526 relies on it)
    return buf


def _shuffle_insert(p: P.Particles, ind: int, val: float) -> P.Particles:
    out = P._empty_like(p)
    L.check(L.lib().mcm_shuffle_insert(p.dtype, C.c_void_p(p.ptr), p.N, int(ind), float(val), C.c_void_p(out.ptr)))
    return out


def with_nominal(p, val, joint: bool = False):
    """with_nominal(p, val) (nominal.jl:20-29): the particle closest to `val` is replaced by `val` and moved to
    index 1 (here: 0); `nominal(pn)` reads it back.

    Vector form: the reference computes `argmin(vec(sum(abs2, M .- val', dims=1)))` with M the N x d sample matrix,
    i.e. the index (among 1..d) of the VARIABLE with the smallest summed squared distance, and uses it as the
    particle index for every variable.  That literal behaviour is the default (drop-in); `joint=True` picks the
    particle nearest to `val` in the joint space, which is what the docstring describes."""
    if isinstance(p, P.Particles):
        if not isinstance(val, numbers.Real):
            raise L.MethodError(L.MCM_ERR_INVALID, "with_nominal: scalar particles need a scalar nominal value")
        ind = C.c_uint64(0)
        L.check(L.lib().mcm_argmin_absdiff(p.dtype, C.c_void_p(p.ptr), p.N, float(val), C.byref(ind)))
        return _shuffle_insert(p, ind.value, float(val))
    ps = list(p)
    vals = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel())
    if len(ps) == 0 or len(vals) != len(ps) or not all(isinstance(q, P.Particles) for q in ps):
        raise L.MethodError(L.MCM_ERR_INVALID, "with_nominal: need one nominal value per particle variable")
    N, dt = ps[0].N, ps[0].dtype
    if any(q.N != N or q.dtype != dt for q in ps):
        raise L.MethodError(L.MCM_ERR_INVALID, "with_nominal: all variables need the same N and element type")
    if joint:
        ptrs = (C.c_void_p * len(ps))(*[C.c_void_p(q.ptr) for q in ps])
        ind = C.c_uint64(0)
        L.check(L.lib().mcm_argmin_sqdist(dt, ptrs, len(ps), N, vals.ctypes.data_as(L._dp), C.byref(ind)))
        k = int(ind.value)
    else:
        # sum_i (x_ij - v_j)^2 = m2_j + N (mean_j - v_j)^2 per variable j, from the one-pass moments
        ssq = []
        for q, v in zip(ps, vals):
            m = P.moments(q)
            ssq.append(m.m2 + m.n * (m.mean - v) ** 2)
        k = int(np.argmin(ssq))
        if k >= N:
            raise L.ArgumentError(L.MCM_ERR_INVALID, "with_nominal: BoundsError in the reference (variable index used as particle index)")
    return [_shuffle_insert(q, k, float(v)) for q, v in zip(ps, vals)]


def nominal(p):
    """nominal(p) (nominal.jl:37-39): `p.particles[1]`; maps over arrays."""
    if isinstance(p, P.Particles):
        host = np.empty(1, dtype=P._NP[p.dtype])
        L.check(L.lib().mcm_download(host.ctypes.data_as(C.c_void_p), C.c_void_p(p.ptr), host.itemsize))
        return host[0].item()
    if isinstance(p, (list, tuple)):
        return type(p)(nominal(q) for q in p)
    if isinstance(p, np.ndarray) and p.dtype == object:
        return np.array([nominal(q) for q in p.ravel()]).reshape(p.shape)
    return p
