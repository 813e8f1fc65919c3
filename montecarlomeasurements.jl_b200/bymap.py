"""bymap / bypmap / @prob (src/bymap.jl:53-136) and register_primitive
(src/register_primitive.jl:9-100).

`bymap` is the reference's escape hatch: an arbitrary host closure evaluated
once per particle.  By construction it cannot be a device kernel, so it keeps
its host semantics here — device -> host export of the arguments, an N-iteration
host loop, host -> device upload of the re-assembled result.
It is part of the reference API, not a fallback for any device primitive.
"""
from __future__ import annotations

import numbers

import numpy as np

from . import _lib as L
from . import particles as P
from . import tape as _tape


def _is_parray(a):
    return P.is_particle_array(a)


def bymap(f, *args):
    """bymap(f, args...) (bymap.jl:67-85): call f on per-particle slices, re-assemble scalar / vector /
    matrix outputs into Particles."""
    # indexof_particles (bymap.jl:32-38)
    pargs = [a for a in args if isinstance(a, P.Particles) or _is_parray(a)]
    if not pargs:
        raise L.ArgumentError(L.MCM_ERR_INVALID, "At least one argument should be <: AbstractParticles.")
    Ns = {P.nparticles(a) for a in pargs}
    if len(Ns) != 1:
        raise L.ArgumentError(L.MCM_ERR_INVALID, "All p::Particles must have the same number of particles.")
    N = Ns.pop()
    # arggetter / vecindex (bymap.jl:24-45): materialise each particle argument once on the host
    host = []
    for a in args:
        if isinstance(a, P.Particles):
            host.append(("p", a.to_numpy()))
        elif _is_parray(a):
            shape = a.shape if isinstance(a, np.ndarray) else (len(a),)
            flat = P.flat_particles(a)
            host.append(("a", np.stack([q.to_numpy() for q in flat], axis=-1).reshape((N,) + tuple(shape))))
        else:
            host.append(("c", a))
    individuals = []
    for i in range(N):
        argsi = [h if k == "c" else h[i] for k, h in host]
        individuals.append(np.asarray(f(*argsi)))
    first = individuals[0]
    stacked = np.stack(individuals, axis=0)
    if stacked.dtype.kind in "iub":
        stacked = stacked.astype(np.float64)
    if first.ndim == 0:
        return P.Particles(np.ascontiguousarray(stacked))
    if first.ndim == 1:
        return [P.Particles(np.ascontiguousarray(stacked[:, j])) for j in range(first.shape[0])]
    if first.ndim == 2:
        out = np.empty(first.shape, dtype=object)
        for r in range(first.shape[0]):
            for c in range(first.shape[1]):
                out[r, c] = P.Particles(np.ascontiguousarray(stacked[:, r, c]))
        return out
    raise RuntimeError("Output with dimension >2 is currently not supported by `bymap`.")  # bymap.jl:83


def bypmap(f, *args):
    """bypmap (bymap.jl:90-108): documented as distributed, but its body is the same `map` as bymap."""
    return bymap(f, *args)


def prob(p: P.Particles, op: str, a: float) -> float:
    """@prob p < a (bymap.jl:128-136) = mean(p.particles .< a); the count runs on device."""
    table = {"<": "LT", "<=": "LE", ">": "GT", ">=": "GE", "==": "EQ", "!=": "NE"}
    return P._count(table[op], p, float(a)) / p.N


# ---- register_primitive ---------------------------------------------------------------------------
_REGISTRY = {}


def register_primitive(f, name: str | None = None):
    """register_primitive(f) (register_primitive.jl:9-12): make `f` work on Particles.

    The reference `eval`s broadcast methods for f.  Here f must be expressible with the registered
    device primitives: it is traced once per arity on symbolic particles and every call becomes ONE
    fused mcm_chain launch.  A function that cannot be traced (arbitrary host code) is exactly the case
    the reference documents `bymap` for — a TapeError says so.
    """
    name = name or getattr(f, "__name__", "primitive")
    fz = P.fused(f)

    def wrapped(*args):
        if not any(isinstance(a, P.Particles) for a in args):
            return f(*args)
        ps = [a for a in args if isinstance(a, P.Particles)]
        key = tuple(isinstance(a, P.Particles) for a in args)
        consts = [a for a in args if not isinstance(a, P.Particles)]
        if consts:
            # close over the scalar arguments: f(p, a::Real...) form (register_primitive.jl:44-47)
            def g(*sym, _args=args):
                it = iter(sym)
                return f(*[next(it) if isinstance(a, P.Particles) else a for a in _args])
            return P.fused(g)(*ps)
        return fz(*ps)

    wrapped.__name__ = name
    _REGISTRY[name] = wrapped
    return wrapped


register_primitive_multi = register_primitive    # register_primitive.jl:19
register_primitive_single = register_primitive   # register_primitive.jl:89
