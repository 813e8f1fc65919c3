"""Device-resident `Particles{T,N}` and the registered primitives.

Host-side mirror of the reference's operator interface for the hot path
(src/types.jl:4-133, src/register_primitive.jl, src/particles.jl:183-190,
302-323, 325-444, 496-593).  The N samples live in HBM behind a raw device
pointer owned by the object (the `particles::Vector{T}` field becomes a
handle); every method body is one call into libmcm_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
import numbers

import numpy as np

from . import _lib as L
from . import tape as _tape

E = L.ENUMS
DEFAULT_NUM_PARTICLES = 2000         # src/MonteCarloMeasurements.jl:47
DEFAULT_STATIC_NUM_PARTICLES = 100   # src/MonteCarloMeasurements.jl:48

_DT = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64}
_NP = {L.F32: np.dtype(np.float32), L.F64: np.dtype(np.float64)}


class ComparisonError(RuntimeError):
    """ErrorException of src/particles.jl:524-533 (comparison of uncertain values failed)."""


class DomainError(ValueError):
    """Raised by check_domain() when a kernel produced NaN from non-NaN input (Base throws DomainError)."""


def _dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _DT:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, f"device particles support Float32/Float64, got {dt}")
    return _DT[dt]


class Particles:
    """`Particles{T,N}` (src/types.jl:17-19) with the samples resident in HBM.

    Constructors (src/types.jl:34-115):
      Particles(v)                  from a 1-D array (types.jl:55)
      Particles(N, dist)            systematic sample of a univariate distribution (types.jl:67-82)
      Particles(N) / Particles()    Normal(0,1), N = 2000 by default
      Particles(T, N, dist)         element type first (types.jl:85-88): ArgumentError unless eltype(dist) == T
      Particles(m) with m 2-D       -> list of Particles, one per column; dim 1 is the particle dim (types.jl:35-52)
    """

    __slots__ = ("ptr", "N", "dtype", "_own", "__weakref__")
    __array_priority__ = 1000

    def __new__(cls, *args, **kw):
        from . import sampling as _s
        if "_ptr" in kw:
            return object.__new__(cls)
        if cls is Particles and len(args) >= 1 and isinstance(args[0], np.ndarray) and args[0].ndim >= 2:
            m = args[0]
            return _from_matrix(m)
        if cls is Particles and len(args) >= 1 and isinstance(args[0], type) and args[0] in (float, np.float32, np.float64):
            return _s.particles_of_type(*args, **kw)                 # PT(::Type{T}, N, d) (types.jl:85-88)
        if cls is Particles and (len(args) == 0 or isinstance(args[0], numbers.Integral) or _s.is_distribution(args[0])):
            return _s.particles_from_distribution(*args, **kw)
        return object.__new__(cls)

    def __init__(self, data=None, *args, _ptr=None, _N=None, _dtype=None, dtype=None, **kw):
        if _ptr is not None:
            self.ptr, self.N, self.dtype, self._own = int(_ptr), int(_N), int(_dtype), True
            return
        if hasattr(self, "ptr"):  # produced by __new__ through a factory
            return
        arr = np.asarray(data)
        if dtype is not None:
            arr = arr.astype(dtype)
        if arr.dtype.kind in "iub":  # Int particles are host-only in the reference; promote like float(p)
            arr = arr.astype(np.float64)
        if arr.ndim != 1:
            raise L.ArgumentError(L.MCM_ERR_INVALID, "Particles(v) expects a vector")
        arr = np.ascontiguousarray(arr)
        self.dtype = _dtype_code(arr.dtype)
        self.N = int(arr.shape[0])
        self.ptr = _alloc(arr.nbytes)
        self._own = True
        L.check(L.lib().mcm_upload(C.c_void_p(self.ptr), arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        L.check(L.lib().mcm_sync())  # the numpy buffer may be a temporary

    def __del__(self):
        try:
            if getattr(self, "_own", False) and getattr(self, "ptr", 0) and L._lib is not None:
                L._lib.mcm_free(C.c_void_p(self.ptr))
        except Exception:
            pass

    # ---- export: Array(p) / Vector(p) (particles.jl:458-459) ----
    @property
    def np_dtype(self):
        return _NP[self.dtype]

    def to_numpy(self) -> np.ndarray:
        out = np.empty(self.N, dtype=self.np_dtype)
        if self.N == 0:
            return out
        L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), out.nbytes))
        return out

    @property
    def particles(self) -> np.ndarray:
        """Materialise the `particles` field on the host (H6: storage is an exposed field in the reference)."""
        return self.to_numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return 1  # a Particles is a Real: length(p) == 1 (bymap.jl:3)

    def __repr__(self):
        try:
            m, s = pmean(self), pstd(self)
            return f"Particles{{{'Float64' if self.dtype == L.F64 else 'Float32'},{self.N}}}({m:.6g} ± {s:.3g})"
        except Exception:
            return f"Particles{{N={self.N}}}(<device {self.ptr:#x}>)"

    def astype(self, dt):
        code = _dtype_code(dt)
        if code == self.dtype:
            return self
        out = _empty_like(self, code)
        L.check(L.lib().mcm_convert(self.dtype, C.c_void_p(self.ptr), code, C.c_void_p(out.ptr), self.N))
        return out

    # ---- operators: register_primitive_multi bodies (register_primitive.jl:44-58) ----
    def __add__(self, o): return _binary("ADD", self, o)
    def __radd__(self, o): return _binary("ADD", o, self)
    def __sub__(self, o): return _binary("SUB", self, o)
    def __rsub__(self, o): return _binary("SUB", o, self)
    def __mul__(self, o): return _binary("MUL", self, o)
    def __rmul__(self, o): return _binary("MUL", o, self)
    def __truediv__(self, o): return _binary("DIV", self, o)
    def __rtruediv__(self, o): return _binary("DIV", o, self)
    def __mod__(self, o): return _binary("MOD", self, o)
    def __rmod__(self, o): return _binary("MOD", o, self)
    # Python's `//` is FLOORED division (pairs with `%` = mod): Julia's fld = div(x, y, RoundDown).
    # Julia's truncating `÷` is div(p, y) below.
    def __floordiv__(self, o): return _binary("DIV_DOWN", self, o)
    def __rfloordiv__(self, o): return _binary("DIV_DOWN", o, self)
    def __neg__(self): return _unary("NEG", self)
    def __pos__(self): return _unary("PLUS", self)
    def __abs__(self): return _unary("ABS", self)

    def __pow__(self, o):
        # Base.:^(p, i::Integer) and literal_pow (particles.jl:383-391)
        if isinstance(o, numbers.Integral) and not isinstance(o, bool):
            out = _empty_like(self)
            L.check(L.lib().mcm_powi(self.dtype, C.c_void_p(self.ptr), int(o), C.c_void_p(out.ptr), self.N))
            return out
        return _binary("POW", self, o)

    def __rpow__(self, o): return _binary("POW", o, self)

    def __round__(self, nd=None):
        if nd not in (None, 0):
            raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "round(p; digits) is not on the device path")
        return _unary("ROUND", self)

    # ---- comparisons (particles.jl:496-546) ----
    def __lt__(self, o): return _compare(self, o, "<")
    def __gt__(self, o): return _compare(o, self, "<")
    def __le__(self, o): return _compare(self, o, "<=")
    def __ge__(self, o): return _compare(o, self, "<=")

    def __eq__(self, o):  # Base.:(==) compares the particle vectors (particles.jl:488)
        if isinstance(o, Particles):
            if o.N != self.N or o.dtype != self.dtype:
                return False
            # all(p1.particles .== p2.particles): one two-input count, no temporary (Inf == Inf, NaN never equal)
            return _count2("EQ", self, o) == self.N
        return NotImplemented

    __hash__ = object.__hash__

    def __bool__(self):
        raise ComparisonError("a Particles value has no truth value; use pmean/≈/≲ or unsafe_comparisons")

    def __float__(self):
        # convert(Float64, p) (particles.jl:334-338): only if all particles are the same
        if self.N == 1:
            return float(self.to_numpy()[0])
        if not pstd(self) < np.finfo(np.float64).eps:
            raise L.ArgumentError(L.MCM_ERR_INVALID,
                                  "Cannot convert a particle distribution to a float if not all particles are the same.")
        return float(pmean(self))


# --------------------------------------------------------------------------------------
def _alloc(nbytes: int) -> int:
    p = C.c_void_p()
    L.check(L.lib().mcm_alloc(C.byref(p), max(int(nbytes), 16)))
    return int(p.value)


def _empty(N: int, dtype: int) -> Particles:
    ptr = _alloc(N * _NP[dtype].itemsize)
    return Particles(_ptr=ptr, _N=N, _dtype=dtype)


def _empty_many(M: int, N: int, dtype: int):
    """M uninitialised Particles of length N: the allocations in ONE C call (mcm_alloc_many) and the handles built
    without the keyword-argument constructor — 1024 separate `_empty` calls cost 4.8 ms of interpreter time, as much
    as the gather that fills them (C5 bootstrap).  One big slab instead was measured too: the pool cannot reuse an
    8 GB block reliably (9-18 ms per call, 150 ms when it has to grow)."""
    if M <= 4:
        return [_empty(N, dtype) for _ in range(M)]
    ptrs = (C.c_void_p * M)()
    L.check(L.lib().mcm_alloc_many(ptrs, M, max(N * _NP[dtype].itemsize, 16)))
    out = []
    new = object.__new__
    for j in range(M):
        q = new(Particles)
        q.ptr = ptrs[j]; q.N = N; q.dtype = dtype; q._own = True
        out.append(q)
    return out


def _empty_like(p: Particles, dtype: int | None = None) -> Particles:
    return _empty(p.N, p.dtype if dtype is None else dtype)


def _from_matrix(m: np.ndarray):
    """PT(m::AbstractArray{T,D}) (types.jl:43-51): dim 1 = particles; returns an array of Particles."""
    m = np.asarray(m)
    tail = m.shape[1:]
    cols = m.reshape(m.shape[0], -1)
    ps = [Particles(np.ascontiguousarray(cols[:, j])) for j in range(cols.shape[1])]
    if len(tail) == 1:
        return ps
    out = np.empty(tail, dtype=object)
    for j, p in enumerate(ps):
        out[np.unravel_index(j, tail)] = p
    return out


def fill(value: float, N: int = DEFAULT_NUM_PARTICLES, dtype=np.float64) -> Particles:
    """PT{T,N}(n::Real) = fill(n, N) (types.jl:57-63) — also what convert(PT, ::Real) gives (particles.jl:332)."""
    code = _dtype_code(dtype)
    out = _empty(int(N), code)
    L.check(L.lib().mcm_fill(code, C.c_void_p(out.ptr), out.N, float(value)))
    return out


def _promote(a: Particles, b: Particles):
    """promote_rule(PT{S,N}, PT{T,N}) = PT{promote_type(S,T),N} (particles.jl:432-440); N must match (MethodError)."""
    if a.N != b.N:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED,
                            f"no method matching op(::Particles{{T,{a.N}}}, ::Particles{{T,{b.N}}}): N must match")
    if a.dtype == b.dtype:
        return a, b
    return a.astype(np.float64), b.astype(np.float64)


def _unary(name: str, p):
    if isinstance(p, _tape.Expr):
        return _tape._un(name, p)
    if not isinstance(p, Particles):
        return _HOST_UNARY[name](p)
    out = _empty_like(p)
    L.check(L.lib().mcm_unary(E["MCM_U_" + name], p.dtype, C.c_void_p(p.ptr), C.c_void_p(out.ptr), p.N))
    return out


def _binary(name: str, a, b):
    if isinstance(a, _tape.Expr) or isinstance(b, _tape.Expr):
        return _tape._bin(name, a, b)
    pa, pb = isinstance(a, Particles), isinstance(b, Particles)
    if not (pa or pb):
        return NotImplemented
    if pa and pb:
        a, b = _promote(a, b)
        out = _empty_like(a)
        L.check(L.lib().mcm_binary(E["MCM_B_" + name], a.dtype, C.c_void_p(a.ptr), 0.0, C.c_void_p(b.ptr), 0.0,
                                   C.c_void_p(out.ptr), a.N))
        return out
    p, s = (a, b) if pa else (b, a)
    if not isinstance(s, numbers.Real):
        return NotImplemented
    # promote_type(Float32 particles, Float64 scalar) = Float64 in Julia only for non-literal Float64 scalars;
    # Python floats are Float64: keep the particle eltype for Float32 (as `p*2` with an Int would).
    out = _empty_like(p)
    op = E["MCM_B_" + name]
    if pa:
        L.check(L.lib().mcm_binary(op, p.dtype, C.c_void_p(p.ptr), 0.0, None, float(s), C.c_void_p(out.ptr), p.N))
    else:
        L.check(L.lib().mcm_binary(op, p.dtype, None, float(s), C.c_void_p(p.ptr), 0.0, C.c_void_p(out.ptr), p.N))
    return out


def _ternary(name: str, a, b, c):
    if any(isinstance(x, _tape.Expr) for x in (a, b, c)):
        return _tape._ter(name, a, b, c)
    ps = [x for x in (a, b, c) if isinstance(x, Particles)]
    if not ps:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "at least one argument must be Particles")
    N = ps[0].N
    if any(p.N != N for p in ps):
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "all Particles must have the same N")
    dt = L.F64 if any(p.dtype == L.F64 for p in ps) else L.F32
    conv = [x.astype(_NP[dt]) if isinstance(x, Particles) else x for x in (a, b, c)]
    out = _empty(N, dt)
    args = []
    for x in conv:
        if isinstance(x, Particles):
            args += [C.c_void_p(x.ptr), 0.0]
        else:
            args += [None, float(x)]
    L.check(L.lib().mcm_ternary(E["MCM_T_" + name], dt, *args, C.c_void_p(out.ptr), N))
    return out


def _count(cmp: str, p: Particles, thr: float = 0.0) -> int:
    c = C.c_uint64()
    L.check(L.lib().mcm_count(E["MCM_CMP_" + cmp], p.dtype, C.c_void_p(p.ptr), p.N, float(thr), C.byref(c)))
    return int(c.value)


def _count2(cmp: str, a: Particles, b: Particles) -> int:
    """#{i : a[i] cmp b[i]} in one pass over both vectors (mcm_count2)."""
    a, b = _promote(a, b)
    c = C.c_uint64()
    L.check(L.lib().mcm_count2(E["MCM_CMP_" + cmp], a.dtype, C.c_void_p(a.ptr), C.c_void_p(b.ptr), a.N, C.byref(c)))
    return int(c.value)


# ---- registered primitives as module-level functions ------------------------------------
_HOST_UNARY = {
    "PLUS": lambda x: +x, "NEG": lambda x: -x, "EXP": math.exp, "EXP2": lambda x: 2.0 ** x, "EXP10": lambda x: 10.0 ** x,
    "EXPM1": math.expm1, "LOG": math.log, "LOG10": math.log10, "LOG2": math.log2, "LOG1P": math.log1p,
    "SIN": math.sin, "COS": math.cos, "TAN": math.tan, "SIND": lambda x: math.sin(math.radians(x)),
    "COSD": lambda x: math.cos(math.radians(x)), "TAND": lambda x: math.tan(math.radians(x)),
    "SINH": math.sinh, "COSH": math.cosh, "TANH": math.tanh, "ASIN": math.asin, "ACOS": math.acos, "ATAN": math.atan,
    "ASIND": lambda x: math.degrees(math.asin(x)), "ACOSD": lambda x: math.degrees(math.acos(x)),
    "ATAND": lambda x: math.degrees(math.atan(x)), "ASINH": math.asinh, "ACOSH": math.acosh, "ATANH": math.atanh,
    "ZERO": lambda x: 0.0, "SIGN": lambda x: (x > 0) - (x < 0), "ABS": abs, "SQRT": math.sqrt,
    "RAD2DEG": math.degrees, "DEG2RAD": math.radians, "FLOAT": float, "ROUND": round, "FLOOR": math.floor,
    "CEIL": math.ceil, "TRUNC": math.trunc, "INV": lambda x: 1.0 / x, "SQR": lambda x: x * x, "CUBE": lambda x: x * x * x,
    "ROUND_TIESAWAY": lambda x: math.copysign(math.floor(abs(x) + 0.5), x),
    "ROUND_TIESUP": lambda x: math.copysign(math.floor(x + 0.5), x),
    "ROUND_FROMZERO": lambda x: math.floor(x) if x < 0 else math.ceil(x),
    "REM2PI_NEAREST": lambda x: math.remainder(x, 2 * math.pi), "REM2PI_DOWN": lambda x: x % (2 * math.pi),
    "REM2PI_TOZERO": lambda x: math.fmod(x, 2 * math.pi), "REM2PI_UP": lambda x: -((-x) % (2 * math.pi)),
}


def _make_unary(name):
    def f(p):
        return _unary(name, p)
    f.__name__ = name.lower()
    f.__doc__ = f"{name.lower()}(p): registered primitive (src/particles.jl:185-190) -> mcm_unary(MCM_U_{name})"
    return f


exp, exp2, exp10, expm1 = (_make_unary(n) for n in ("EXP", "EXP2", "EXP10", "EXPM1"))
log, log10, log2, log1p = (_make_unary(n) for n in ("LOG", "LOG10", "LOG2", "LOG1P"))
sin, cos, tan, sind, cosd, tand = (_make_unary(n) for n in ("SIN", "COS", "TAN", "SIND", "COSD", "TAND"))
sinh, cosh, tanh = (_make_unary(n) for n in ("SINH", "COSH", "TANH"))
asin, acos, asind, acosd = (_make_unary(n) for n in ("ASIN", "ACOS", "ASIND", "ACOSD"))
asinh, acosh, atanh = (_make_unary(n) for n in ("ASINH", "ACOSH", "ATANH"))
zero, sign, sqrt, rad2deg, deg2rad = (_make_unary(n) for n in ("ZERO", "SIGN", "SQRT", "RAD2DEG", "DEG2RAD"))
floor, ceil, trunc, inv, abs2 = (_make_unary(n) for n in ("FLOOR", "CEIL", "TRUNC", "INV", "SQR"))
float_ = _make_unary("FLOAT")


def atan(y, x=None):
    """atan(p) / atan(y, x) — both are registered (particles.jl:183,188)."""
    return _unary("ATAN", y) if x is None else _binary("ATAN2", y, x)


def atand(y, x=None):
    return _unary("ATAND", y) if x is None else _binary("ATAND2", y, x)


# ---- RoundingMode forms (particles.jl:346-348, 358-359, 410-414; runtests.jl:470-483) -------------------------
RoundNearest, RoundToZero, RoundUp, RoundDown = "RoundNearest", "RoundToZero", "RoundUp", "RoundDown"
RoundNearestTiesAway, RoundNearestTiesUp, RoundFromZero = "RoundNearestTiesAway", "RoundNearestTiesUp", "RoundFromZero"
_ROUND_OP = {RoundNearest: "ROUND", RoundToZero: "TRUNC", RoundUp: "CEIL", RoundDown: "FLOOR",
             RoundNearestTiesAway: "ROUND_TIESAWAY", RoundNearestTiesUp: "ROUND_TIESUP", RoundFromZero: "ROUND_FROMZERO"}
_RM_CODE = {RoundNearest: "NEAREST", RoundToZero: "TOZERO", RoundUp: "UP", RoundDown: "DOWN",
            RoundNearestTiesAway: "TIESAWAY", RoundNearestTiesUp: "TIESUP", RoundFromZero: "FROMZERO"}
_REM_OP = {RoundToZero: "REM", RoundNearest: "REM_NEAREST", RoundDown: "REM_DOWN", RoundUp: "REM_UP",
           RoundFromZero: "REM_FROMZERO"}
_DIV_OP = {RoundToZero: "FDIV", RoundNearest: "DIV_NEAREST", RoundDown: "DIV_DOWN", RoundUp: "DIV_UP",
           RoundFromZero: "DIV_FROMZERO"}
_REM2PI_OP = {RoundNearest: "REM2PI_NEAREST", RoundToZero: "REM2PI_TOZERO", RoundDown: "REM2PI_DOWN", RoundUp: "REM2PI_UP"}


def _mode(table, r, what):
    if r not in table:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, f"{what}: unsupported rounding mode {r!r}")
    return table[r]


class IntParticles:
    """`Particles{Int,N}` as produced by round(Int, p[, r]) (particles.jl:347-348): N Int64 samples in HBM.
    Integer particles are host-only in the reference; on device they exist as the result of rounding
    (to_numpy / comparison with an integer / float(q) -> Particles{Float64})."""
    __slots__ = ("index", "N")

    def __init__(self, index: "DeviceIndex"):
        self.index, self.N = index, index.n

    def to_numpy(self) -> np.ndarray:
        return self.index.to_numpy()

    @property
    def particles(self):
        return self.to_numpy()

    def __eq__(self, o):
        if isinstance(o, numbers.Integral):       # p == 0 (runtests.jl:471): every particle equals o
            return bool(np.all(self.to_numpy() == int(o)))
        if isinstance(o, IntParticles):
            return o.N == self.N and bool(np.array_equal(self.to_numpy(), o.to_numpy()))
        return NotImplemented

    __hash__ = object.__hash__

    def float(self) -> "Particles":
        return Particles(self.to_numpy().astype(np.float64))


def round_(p, r=RoundNearest):
    """round(p) / round(p, r::RoundingMode) (particles.jl:346)."""
    return _unary(_mode(_ROUND_OP, r, "round"), p)


def round_int(p: Particles, r=RoundNearest) -> IntParticles:
    """round(Int, p[, r]) (particles.jl:347-348): Int64 particles; NaN / Inf / overflow raise DomainError at
    check_domain() (InexactError in Base)."""
    d = DeviceIndex(p.N)
    L.check(L.lib().mcm_round_int(p.dtype, C.c_void_p(p.ptr), E["MCM_RM_" + _mode(_RM_CODE, r, "round(Int, p, r)")],
                                  C.c_void_p(d.ptr), p.N))
    return IntParticles(d)


def rem2pi(p, r=RoundNearest):
    """rem2pi(p, r) (particles.jl:359): correctly rounded remainder modulo 2pi for every finite argument."""
    return _unary(_mode(_REM2PI_OP, r, "rem2pi"), p)


def mod2pi(p):
    """mod2pi(p) = rem2pi(p, RoundDown) (runtests.jl:476)."""
    return _unary("REM2PI_DOWN", p)


def isinteger(p: Particles) -> bool:
    """all(isinteger, p.particles) (particles.jl:583)."""
    return _count("ISINTEGER", p) == p.N


def iszero(p: Particles, tol=None) -> bool:
    """iszero(p) = all(iszero, p.particles); iszero(p, tol) = abs(mean(p.particles)) < tol (particles.jl:584-585)."""
    if tol is not None:
        return abs(pmean(p)) < tol
    return _count("EQ", p, 0.0) == p.N


def not_(p: Particles) -> bool:
    """!p = all(p.particles .== 0) (particles.jl:581)."""
    return _count("EQ", p, 0.0) == p.N


def isfinite(p: Particles) -> bool:
    """isfinite(p) = isfinite(pmean(p)) (particles.jl:345)."""
    return math.isfinite(pmean(p))


def maximum(a, b): return _binary("MAX", a, b)   # Base.max(p1, p2)
def minimum(a, b): return _binary("MIN", a, b)   # Base.min(p1, p2)
def mod(a, b): return _binary("MOD", a, b)
def mod1(a, b): return _binary("MOD1", a, b)
def rem(a, b, r=RoundToZero):
    """rem(p, y) / rem(p, y, r::RoundingMode) (particles.jl:354-355; Base div.jl for the modes)."""
    return _binary(_mode(_REM_OP, r, "rem"), a, b)


def div(a, b, r=RoundToZero):
    """div(p, y) = round((x - rem(x, y))/y) and div(p, y, r) (particles.jl:356-358, 410-414)."""
    return _binary(_mode(_DIV_OP, r, "div"), a, b)


def fld(a, b): return _binary("DIV_DOWN", a, b)   # floored division (Python's //)
def cld(a, b): return _binary("DIV_UP", a, b)
def hypot(a, b): return _binary("HYPOT", a, b)
def add_sum(a, b): return _binary("ADD_SUM", a, b)
def clamp(x, lo, hi): return _ternary("CLAMP", x, lo, hi)
def muladd(x, y, z): return _ternary("MULADD", x, y, z)   # particles.jl:416-422
def minmax(x, y): return minimum(x, y), maximum(x, y)     # particles.jl:579
def sincos(x): return sin(x), cos(x)                       # particles.jl:578


def ratdiv(a, b):
    """`//` is in the registered list (particles.jl:183) but has no method for floats in Base."""
    pa = a if isinstance(a, Particles) else b
    out = _empty_like(pa)
    L.check(L.lib().mcm_binary(E["MCM_B_RATDIV"], pa.dtype, C.c_void_p(pa.ptr), 0.0, C.c_void_p(pa.ptr), 0.0,
                               C.c_void_p(out.ptr), pa.N))


# ---- fused chains -----------------------------------------------------------------------
def evalpoly(p: Particles, coeffs, fused: bool = True) -> Particles:
    """evalpoly(p, coeffs) (Horner).  fused=True uses muladd like Base.evalpoly; False rounds * and + separately."""
    c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64))
    out = _empty_like(p)
    L.check(L.lib().mcm_polyval(p.dtype, C.c_void_p(p.ptr), c.ctypes.data_as(L._dp), len(c), 1 if fused else 0,
                                C.c_void_p(out.ptr), p.N))
    return out


def affine(p: Particles, alpha: float, beta: float) -> Particles:
    """alpha*p + beta in one pass, bit-identical to the two reference broadcasts."""
    out = _empty_like(p)
    L.check(L.lib().mcm_affine(p.dtype, C.c_void_p(p.ptr), float(alpha), float(beta), C.c_void_p(out.ptr), p.N))
    return out


def lincomb(coeffs, ps) -> Particles:
    """c0*p0 + c1*p1 + ... in one pass (RK4 stage updates), left to right, unfused roundings."""
    ps = list(ps)
    N, dt = ps[0].N, ps[0].dtype
    if any(p.N != N or p.dtype != dt for p in ps):
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "lincomb: all Particles must share N and eltype")
    c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64))
    ptrs = (C.c_void_p * len(ps))(*[C.c_void_p(p.ptr) for p in ps])
    out = _empty(N, dt)
    L.check(L.lib().mcm_lincomb(dt, len(ps), ptrs, c.ctypes.data_as(L._dp), C.c_void_p(out.ptr), N))
    return out


class fused:
    """`fused(f)(p1, p2, ...)`: evaluate a whole expression of registered primitives in ONE kernel.

    f is traced once on symbolic operands (tape.py); scalars are captured as constants.
    With `stats=True` the six moments of the (first) result are reduced in the same pass and
    returned as second value; with `materialize=False` the result vector is never written to HBM.
    """

    def __init__(self, fn):
        self.fn = fn
        self._cache = {}

    def tape(self, nin: int) -> _tape.Tape:
        if nin not in self._cache:
            self._cache[nin] = _tape.trace(self.fn, nin)
        return self._cache[nin]

    def __call__(self, *ps, stats: bool = False, materialize: bool = True):
        if not ps or not all(isinstance(p, Particles) for p in ps):
            raise L.ArgumentError(L.MCM_ERR_INVALID, "fused(f)(...) takes Particles arguments")
        N = ps[0].N
        if any(p.N != N for p in ps):
            raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "all Particles must have the same N")
        dt = L.F64 if any(p.dtype == L.F64 for p in ps) else L.F32
        ps = [p.astype(_NP[dt]) for p in ps]
        t = self.tape(len(ps))
        outs = [_empty(N, dt) for _ in range(t.nout)] if materialize else []
        st = _tape.run_tape(t, [p.ptr for p in ps], N, dt, [o.ptr for o in outs] if materialize else None,
                            stats_output=0 if stats else None)
        res = (outs[0] if t.nout == 1 else tuple(outs)) if materialize else None
        if stats:
            return res, Moments(*st)
        return res


class Moments:
    """{n, mean, M2, min, max, sum} of one particle vector (one HBM pass)."""
    __slots__ = ("n", "mean", "m2", "min", "max", "sum")

    def __init__(self, n, mean, m2, mn, mx, s):
        self.n, self.mean, self.m2, self.min, self.max, self.sum = n, mean, m2, mn, mx, s

    def var(self, corrected=True):
        if self.n == 1:
            return 0.0  # particles.jl:309: N == 1 && return zero(T)
        return self.m2 / (self.n - (1.0 if corrected else 0.0))

    def std(self, corrected=True):
        return math.sqrt(self.var(corrected))


# ---- statistics: particles.jl:302-323 ----------------------------------------------------
def moments(p: Particles) -> Moments:
    m = np.zeros(6)
    L.check(L.lib().mcm_moments(p.dtype, C.c_void_p(p.ptr), p.N, m.ctypes.data_as(L._dp)))
    return Moments(*m)


def _reduce(kind: str, p: Particles) -> float:
    r = C.c_double()
    L.check(L.lib().mcm_reduce(E["MCM_R_" + kind], p.dtype, C.c_void_p(p.ptr), p.N, C.byref(r)))
    return r.value


def _scalar_or(p, f_num, f_part):
    if isinstance(p, Particles):
        return f_part(p)
    if isinstance(p, (list, tuple, np.ndarray)):
        return [(_scalar_or(x, f_num, f_part)) for x in p]
    return f_num(p)  # pmean(p::Number) passthrough (particles.jl:322)


def pmean(p):
    """pmean(p) = mean(p.particles) (particles.jl:321); vectors map elementwise (particles.jl:475)."""
    return _scalar_or(p, float, lambda q: _reduce("MEAN", q))


def psum(p):
    return _scalar_or(p, float, lambda q: _reduce("SUM", q))


def pvar(p, corrected: bool = True, mean=None):
    """pvar(p; corrected, mean) (particles.jl:308-312)."""
    def f(q):
        if q.N == 1:
            return 0.0
        if mean is None:
            return _reduce("VAR" if corrected else "VARP", q)
        # var(x; mean=m): sum((x-m)^2)/(n-corrected)
        d = _binary("SUB", q, float(mean))
        return _reduce("SUM", _unary("SQR", d)) / (q.N - (1 if corrected else 0))
    return _scalar_or(p, lambda x: 0.0, f)


def pstd(p, corrected: bool = True, mean=None):
    v = pvar(p, corrected=corrected, mean=mean)
    return [math.sqrt(x) for x in v] if isinstance(v, list) else math.sqrt(v)


def pminimum(p): return _scalar_or(p, float, lambda q: _reduce("MIN", q))
def pmaximum(p): return _scalar_or(p, float, lambda q: _reduce("MAX", q))


def pextrema(p):
    def f(q):
        m = moments(q)
        return (m.min, m.max)
    return _scalar_or(p, lambda x: (float(x), float(x)), f)


def pmiddle(p):
    """middle(extrema) = (min+max)/2 (Statistics.middle of the particle vector)."""
    lo, hi = pextrema(p)
    return lo / 2 + hi / 2


def piterate(p, state: int | None = None):
    """piterate(p[, state]) = Base.iterate(p.particles[, state]) (particles.jl:316-323): (particle, next state) with a
    1-based state like Julia's, or None past the end; one 8-byte read from the device per step.  A Number iterates once."""
    if not isinstance(p, Particles):
        return (p, None) if state is None else None
    i = 1 if state is None else int(state)
    if i < 1 or i > p.N:
        return None
    out = np.empty(1, dtype=p.np_dtype)
    L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(p.ptr + (i - 1) * out.itemsize), out.itemsize))
    return out[0].item(), i + 1


def meanstd(p): return pstd(p) / math.sqrt(p.N)    # particles.jl:485
def meanvar(p): return pvar(p) / p.N               # particles.jl:486


def pquantile(p, q):
    """pquantile(p, q) = Statistics.quantile(p.particles, q) (particles.jl:321); q scalar or vector.
    p may be a vector of Particles: one batched device call (mcm_quantile_batch), result [M] or [M, nq]."""
    scalar = isinstance(q, numbers.Real)
    qs = np.ascontiguousarray(np.atleast_1d(np.asarray(q, dtype=np.float64)))
    if not isinstance(p, Particles):
        v = _check_mv(p)
        M = len(v)
        ptrs = (C.c_void_p * M)(*[C.c_void_p(x.ptr) for x in v])
        out = np.zeros((M, len(qs)))
        L.check(L.lib().mcm_quantile_batch(v[0].dtype, ptrs, M, v[0].N, qs.ctypes.data_as(L._dp), len(qs),
                                           out.ctypes.data_as(L._dp)))
        return out[:, 0].copy() if scalar else out
    out = np.zeros(len(qs))
    L.check(L.lib().mcm_quantile(p.dtype, C.c_void_p(p.ptr), p.N, qs.ctypes.data_as(L._dp), len(qs),
                                 out.ctypes.data_as(L._dp)))
    return float(out[0]) if scalar else out


def order_statistics(p: Particles, ranks) -> np.ndarray:
    """sort(p.particles)[ranks] (0-based) without sorting (mcm_order_statistics)."""
    rk = np.ascontiguousarray(np.atleast_1d(np.asarray(ranks, dtype=np.uint64)))
    out = np.zeros(len(rk))
    L.check(L.lib().mcm_order_statistics(p.dtype, C.c_void_p(p.ptr), p.N, rk.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         len(rk), out.ctypes.data_as(L._dp)))
    return out


def pmedian(p):
    """Statistics.median: middle of the one/two central order statistics (no full sort)."""
    n = p.N
    if n == 0:
        raise L.ArgumentError(L.MCM_ERR_INVALID, "median of an empty collection")
    if n % 2 == 1:
        return float(order_statistics(p, [n // 2])[0])
    a, b = order_statistics(p, [n // 2 - 1, n // 2])
    if p.np_dtype == np.float32:
        a, b = np.float32(a), np.float32(b)
        return float(a / np.float32(2) + b / np.float32(2))
    return float(a / 2 + b / 2)  # Statistics.middle(x, y) = x/2 + y/2


def check_bounds():
    """Raise IndexError if a gather met an index outside 1:N since the last check (BoundsError in the reference,
    resampling.jl:77,82).  Called after every gather with caller-supplied indices; synchronises."""
    bits = C.c_uint32()
    L.check(L.lib().mcm_check_flags(C.byref(bits), 0))
    if bits.value & 0x80000000:
        L.check(L.lib().mcm_clear_flags(0x80000000))
        raise IndexError("BoundsError: resample/gather index out of range")


def _pick(p: Particles, idx):
    idx_dev = _upload_i64(np.asarray(idx, dtype=np.int64))
    out = _empty(len(idx), p.dtype)
    L.check(L.lib().mcm_gather(p.dtype, C.c_void_p(p.ptr), C.c_void_p(idx_dev.ptr), C.c_void_p(out.ptr), len(idx), p.N))
    check_bounds()
    return out.to_numpy().astype(np.float64)


def psort(p: Particles) -> Particles:
    out = _empty_like(p)
    L.check(L.lib().mcm_sort(p.dtype, C.c_void_p(p.ptr), C.c_void_p(out.ptr), p.N))
    return out


class DeviceIndex:
    """A device-resident int64 index vector (0-based) — permutations, resample indices."""
    __slots__ = ("ptr", "n")

    def __init__(self, n: int, ptr: int | None = None):
        self.n = int(n)
        self.ptr = _alloc(self.n * 8) if ptr is None else int(ptr)

    def __del__(self):
        try:
            if self.ptr and L._lib is not None:
                L._lib.mcm_free(C.c_void_p(self.ptr))
        except Exception:
            pass

    def to_numpy(self):
        out = np.empty(self.n, dtype=np.int64)
        L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), out.nbytes))
        return out


def _upload_i64(a: np.ndarray) -> DeviceIndex:
    a = np.ascontiguousarray(a, dtype=np.int64)
    d = DeviceIndex(len(a))
    L.check(L.lib().mcm_upload(C.c_void_p(d.ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))
    L.check(L.lib().mcm_sync())
    return d


def psortperm(p: Particles) -> DeviceIndex:
    """sortperm(p.particles) as a 0-based device index vector (stable)."""
    d = DeviceIndex(p.N)
    L.check(L.lib().mcm_sortperm(p.dtype, C.c_void_p(p.ptr), C.c_void_p(d.ptr), None, p.N))
    return d


# ---- multivariate statistics: particles.jl:461-465, 475-478 -------------------------------
def _check_mv(v):
    v = list(v)
    if not v or not all(isinstance(p, Particles) for p in v):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "expected a vector of Particles")
    N, dt = v[0].N, v[0].dtype
    if any(p.N != N for p in v):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "All p::Particles must have the same number of particles.")
    if any(p.dtype != dt for p in v):
        v = [p.astype(np.float64) for p in v]
    return v


def to_matrix(v) -> np.ndarray:
    """Matrix(v::MvParticles) (particles.jl:461-465): N x M, column j = variable j (host export)."""
    v = _check_mv(v)
    return np.stack([p.to_numpy() for p in v], axis=1)


def pcov(v, corrected: bool = True) -> np.ndarray:
    """pcov(v::MvParticles) = cov(Matrix(v)) (particles.jl:476).  Scalar p: pcov(p) = var (particles.jl:321)."""
    if isinstance(v, Particles):
        return pvar(v, corrected=corrected)
    v = _check_mv(v)
    M = len(v)
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
    dC = _alloc(M * M * 8)
    try:
        L.check(L.lib().mcm_gram(v[0].dtype, ptrs, M, v[0].N, 1 if corrected else 0, None, C.c_void_p(dC), None))
        out = np.empty((M, M))
        L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(dC), out.nbytes))
    finally:
        L.lib().mcm_free(C.c_void_p(dC))
    return out


# ---- Particle BLAS (src/particles.jl:614-707) ------------------------------------------------
def pgemm(A, xs, shift=None, bias=None):
    """y_j = bias_j + sum_k A[j,k] (x_k - shift_k) on device (mcm_pgemm): the GEMM behind
    mul!(y, A, b::Vector{Particles}) (particles.jl:688-707), A*p (_pgemv, :620-631) and the
    transform_moments application (sigmapoints.jl:71).  Returns a list of Mo new Particles."""
    xs = _check_mv(xs)
    A = np.ascontiguousarray(np.atleast_2d(np.asarray(A, dtype=np.float64)))
    Mo, Mi = A.shape
    if Mi != len(xs):
        raise L.ArgumentError(L.MCM_ERR_INVALID, f"DimensionMismatch: A has {Mi} columns, vector has {len(xs)} entries")
    sh = None if shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(shift, dtype=np.float64), (Mi,)))
    bi = None if bias is None else np.ascontiguousarray(np.broadcast_to(np.asarray(bias, dtype=np.float64), (Mo,)))
    outs = [_empty_like(xs[0]) for _ in range(Mo)]
    xin = (C.c_void_p * Mi)(*[C.c_void_p(x.ptr) for x in xs])
    xout = (C.c_void_p * Mo)(*[C.c_void_p(y.ptr) for y in outs])
    L.check(L.lib().mcm_pgemm(xs[0].dtype, A.ctypes.data_as(L._dp), Mo, Mi, xin,
                              None if sh is None else sh.ctypes.data_as(L._dp),
                              None if bi is None else bi.ctypes.data_as(L._dp), xout, xs[0].N))
    return outs


def pgemv(A, p):
    """A * p::Vector{Particles} (particles.jl:620-633)."""
    return pgemm(A, p)


def mul_(y, A, b):
    """mul!(y, A, b) (particles.jl:688-707): fills the list y in place with A*b and returns it."""
    res = pgemm(A, b)
    if len(y) != len(res):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "DimensionMismatch: y has the wrong length")
    y[:] = res
    return y


def pdot(v, p):
    """dot(v, p::Vector{Particles}) (_pdot, particles.jl:641-654)."""
    return pgemm(np.asarray(v, dtype=np.float64)[None, :], p)[0]


def paxpy_(a, x, y):
    """axpy!(a, x, y) on particle vectors (_paxpy!, particles.jl:657-673): y .= a*x + y, returns y."""
    x, yv = _check_mv(x), _check_mv(y)
    if len(x) != len(yv):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "DimensionMismatch")
    y[:] = [muladd(float(a), xi, yi) for xi, yi in zip(x, yv)]
    return y


def pcor(v) -> np.ndarray:
    """pcor(v) = cor(Matrix(v)) (particles.jl:477): covariance normalised by sqrt of the diagonal."""
    c = pcov(v)
    d = np.sqrt(np.diag(c))
    r = c / d[:, None] / d[None, :]
    np.fill_diagonal(r, 1.0)
    return np.clip(r, -1.0, 1.0)  # Statistics.cor clamps to [-1,1]


# ---- comparisons (src/MonteCarloMeasurements.jl:55-118, src/particles.jl:496-546) ----------
_COMPARISON_MODE = ["safe"]
_COMPARISON_FUNCTION = [pmean]


def unsafe_comparisons(mode=True):
    """unsafe_comparisons(mode): :safe / :montecarlo / :reduction (MonteCarloMeasurements.jl:67-79)."""
    if mode is True: mode = "reduction"
    if mode is False: mode = "safe"
    if mode not in ("safe", "montecarlo", "reduction"):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "Got unsupported comparison mode.")
    _COMPARISON_MODE[0] = mode


def set_comparison_function(f):
    _COMPARISON_FUNCTION[0] = f


def _compare(a, b, op: str) -> bool:
    mode = _COMPARISON_MODE[0]
    f = (lambda x, y: x < y) if op == "<" else (lambda x, y: x <= y)
    if mode == "reduction":
        r = _COMPARISON_FUNCTION[0]
        return f(r(a) if isinstance(a, Particles) else a, r(b) if isinstance(b, Particles) else b)
    if mode == "safe":  # particles.jl:502-507: all / none over extrema x extrema
        ea = pextrema(a) if isinstance(a, Particles) else (a, a)
        eb = pextrema(b) if isinstance(b, Particles) else (b, b)
        res = [f(x, y) for x in ea for y in eb]
        if all(res): return True
        if not any(res): return False
        raise ComparisonError("Comparison of uncertain values using comparison mode safe failed.")
    # montecarlo: elementwise all / none (particles.jl:512-515)
    cmp = "LT" if op == "<" else "LE"
    if isinstance(a, Particles) and isinstance(b, Particles):
        N, c = a.N, _count2(cmp, a, b)
    elif isinstance(a, Particles):
        N, c = a.N, _count(cmp, a, float(b))
    else:
        N, c = b.N, _count("GT" if op == "<" else "GE", b, float(a))
    if c == N: return True
    if c == 0: return False
    raise ComparisonError("Comparison of uncertain values using comparison mode montecarlo failed.")


def approx(p, a, lim: float = 2.0) -> bool:
    """p ≈ a (particles.jl:553-568)."""
    if isinstance(p, Particles) and isinstance(a, Particles):
        return abs(pmean(p) - pmean(a)) / (2 * math.sqrt(pstd(p) ** 2 + pstd(a) ** 2)) < lim
    if isinstance(a, Particles):
        p, a = a, p
    m = pmean(p)
    s = pstd(p, mean=m)
    if s == 0:
        return m == a
    return abs(m - a) / pstd(p) < lim


def lesssim(a, b, lim: float = 2.0) -> bool:
    """a ≲ b (particles.jl:587-590)."""
    if isinstance(a, Particles) and isinstance(b, Particles):
        return (pmean(b) - pmean(a)) / (2 * math.sqrt(pstd(a) ** 2 + pstd(b) ** 2)) > lim
    if isinstance(b, Particles):
        return (pmean(b) - a) / pstd(b) > lim
    if isinstance(a, Particles):
        return (b - pmean(a)) / pstd(a) > lim
    return a < b


def gtrsim(a, b, lim: float = 2.0) -> bool:
    return lesssim(b, a, lim)


def check_domain(reset: bool = True):
    """Surface the sticky DOMAIN flag as the DomainError Base would have thrown (H7)."""
    bits = C.c_uint32()
    L.check(L.lib().mcm_check_flags(C.byref(bits), 1 if reset else 0))
    if bits.value & E["MCM_FLAG_DOMAIN"]:
        raise DomainError("a primitive was evaluated outside its domain (NaN produced from non-NaN input)")
    if bits.value & 0x80000000:
        raise IndexError("BoundsError: resample/gather index out of range")
    return bits.value


def flat_particles(p):
    """Elements of a Particles / list / tuple / object ndarray of Particles, in storage order.
    (np.asarray on a list of Particles would call __array__ and download them — never do that.)"""
    if isinstance(p, Particles):
        return [p]
    if isinstance(p, np.ndarray):
        return list(p.flat)
    out = []
    for q in p:
        out.extend(flat_particles(q) if isinstance(q, (list, tuple, np.ndarray)) else [q])
    return out


def is_particle_array(a) -> bool:
    if isinstance(a, np.ndarray):
        return a.dtype == object and a.size > 0 and isinstance(a.flat[0], Particles)
    return isinstance(a, (list, tuple)) and len(a) > 0 and all(isinstance(q, Particles) for q in flat_particles(a))


def nparticles(p) -> int:
    """nparticles (bymap.jl:3-8)."""
    if isinstance(p, Particles): return p.N
    if is_particle_array(p): return flat_particles(p)[0].N
    return 1


def particleeltype(p):
    """particleeltype (bymap.jl:14-15)."""
    if isinstance(p, Particles): return p.np_dtype
    return flat_particles(p)[0].np_dtype
