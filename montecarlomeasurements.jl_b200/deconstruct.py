"""has_particles / replace_particles / mean_object (src/deconstruct.jl:1-20, 90-130): host-side traversal of an object
that carries uncertain fields.  The reference walks `fieldnames(typeof(P))`; the mirror walks lists, tuples, dicts,
numpy object arrays, dataclasses and plain objects with a `__dict__`.  The statistics called on the leaves (`pmean`)
run on the device; the traversal itself is host logic by nature (like `bymap`)."""
from __future__ import annotations

import copy
import dataclasses
import numbers

import numpy as np

from . import particles as P


def _is_leaf(x) -> bool:
    from .complexp import ComplexParticles
    return isinstance(x, (P.Particles, ComplexParticles))


def _children(x):
    if isinstance(x, dict):
        return list(x.values())
    if isinstance(x, (list, tuple)):
        return list(x)
    if isinstance(x, np.ndarray) and x.dtype == object:
        return list(x.ravel())
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        return [getattr(x, f.name) for f in dataclasses.fields(x)]
    if hasattr(x, "__dict__") and not isinstance(x, type):
        return list(vars(x).values())
    return []


def has_particles(x) -> bool:
    """deconstruct.jl:5-20: does the object hold particles anywhere inside (fields, arrays, tuples)?"""
    if _is_leaf(x):
        return True
    if isinstance(x, (numbers.Number, str, bytes)) or (isinstance(x, np.ndarray) and x.dtype != object):
        return False
    return any(has_particles(c) for c in _children(x))


def replace_particles(x, condition=None, replacer=None):
    """deconstruct.jl:102-130: rebuild `x` with every field that satisfies `condition` replaced by `replacer(field)`
    (defaults: any particles -> their first sample, `vecindex(P, 1)`)."""
    from .complexp import ComplexParticles
    condition = condition or _is_leaf
    replacer = replacer or (lambda p: (p.re.to_numpy()[0] + 1j * p.im.to_numpy()[0]) if isinstance(p, ComplexParticles) else p.to_numpy()[0])
    if condition(x):
        return replacer(x)
    if not has_particles(x):
        return x
    rec = lambda c: replace_particles(c, condition, replacer)
    if isinstance(x, dict):
        return {k: rec(v) for k, v in x.items()}
    if isinstance(x, tuple):
        vals = [rec(v) for v in x]
        return type(x)(*vals) if hasattr(x, "_fields") else tuple(vals)       # namedtuple / tuple
    if isinstance(x, list):
        return [rec(v) for v in x]
    if isinstance(x, np.ndarray):
        out = np.empty(x.shape, dtype=object)
        out.ravel()[:] = [rec(v) for v in x.ravel()]
        try:
            return out.astype(np.float64)
        except (TypeError, ValueError):
            return out
    if dataclasses.is_dataclass(x):
        return dataclasses.replace(x, **{f.name: rec(getattr(x, f.name)) for f in dataclasses.fields(x)})
    y = copy.copy(x)
    for k, v in vars(x).items():
        setattr(y, k, rec(v))
    return y


def mean_object(x):
    """deconstruct.jl:93-95: the same object with every `Particles` replaced by its mean."""
    from .complexp import ComplexParticles
    return replace_particles(x, replacer=lambda p: complex(P.pmean(p.re), P.pmean(p.im)) if isinstance(p, ComplexParticles) else P.pmean(p))
