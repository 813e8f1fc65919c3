"""Complex{Particles} (src/complex.jl): an uncertain complex number is a pair of device-resident Particles.

Reference behaviour mirrored here: `+ - *` are Base's generic Complex arithmetic on the two Particles fields; `/`, `^`
and `sqrt exp exp10 log log10 sin cos tan` go through `ℂ2ℂ_function` (complex.jl:30-80), which loops over the particle
index on the host; statistics act componentwise (complex.jl:96-98).  On device every operation is one kernel
(`mcm_complex_op`)."""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from . import _lib as L
from . import particles as P

E = L.ENUMS


class ComplexParticles:
    """`Complex{Particles{T,N}}`: `re` and `im` are Particles with the same N and element type."""
    __slots__ = ("re", "im")
    __array_priority__ = 1001

    def __init__(self, re, im=None):
        if isinstance(re, ComplexParticles):
            re, im = re.re, re.im
        if not isinstance(re, P.Particles):
            arr = np.asarray(re)
            if np.iscomplexobj(arr):                      # from complex samples (switch_representation, complex.jl:100)
                re, im = P.Particles(np.ascontiguousarray(arr.real)), P.Particles(np.ascontiguousarray(arr.imag))
            else:
                re = P.Particles(arr)
        if im is None:
            im = P.fill(0.0, re.N, P._NP[re.dtype])
        elif isinstance(im, numbers.Real):
            im = P.fill(float(im), re.N, P._NP[re.dtype])
        elif not isinstance(im, P.Particles):
            im = P.Particles(np.asarray(im))
        if re.N != im.N or re.dtype != im.dtype:
            raise L.MethodError(L.MCM_ERR_INVALID, "Complex{Particles}: real and imaginary parts need the same N and element type")
        self.re, self.im = re, im

    # ---- plumbing -------------------------------------------------------------------------------
    @property
    def N(self): return self.re.N
    @property
    def dtype(self): return self.re.dtype

    def to_numpy(self) -> np.ndarray:
        return self.re.to_numpy() + 1j * self.im.to_numpy()

    def __repr__(self):
        return f"Complex({self.re!r}, {self.im!r})"

    @staticmethod
    def _operand(x, like: "ComplexParticles"):
        """-> (re_ptr|None, im_ptr|None, sre, sim) for the C ABI; checks N / element type."""
        if isinstance(x, ComplexParticles):
            if x.N != like.N or x.dtype != like.dtype:
                raise L.MethodError(L.MCM_ERR_INVALID, "Complex{Particles}: operands need the same N and element type")
            return C.c_void_p(x.re.ptr), C.c_void_p(x.im.ptr), 0.0, 0.0, (x.re, x.im)
        if isinstance(x, P.Particles):
            if x.N != like.N or x.dtype != like.dtype:
                raise L.MethodError(L.MCM_ERR_INVALID, "Complex{Particles}: operands need the same N and element type")
            return C.c_void_p(x.ptr), None, 0.0, 0.0, (x,)
        if isinstance(x, numbers.Complex):
            z = complex(x)
            return None, None, z.real, z.imag, ()
        raise L.MethodError(L.MCM_ERR_INVALID, f"Complex{{Particles}}: unsupported operand {type(x).__name__}")

    def _apply(self, op: str, a, b=None):
        are, aim, sar, sai, keep_a = self._operand(a, self)
        bre, bim, sbr, sbi, keep_b = self._operand(b, self) if b is not None else (None, None, 0.0, 0.0, ())
        ore, oim = P._empty_like(self.re), P._empty_like(self.re)
        L.check(L.lib().mcm_complex_op(E["MCM_C_" + op], self.dtype, are, aim, sar, sai, bre, bim, sbr, sbi,
                                       C.c_void_p(ore.ptr), C.c_void_p(oim.ptr), self.N))
        del keep_a, keep_b
        return ComplexParticles(ore, oim)

    # ---- arithmetic (Base's Complex methods on the Particles fields; complex.jl:7-16, :85-91) ----
    def __add__(self, o): return self._apply("ADD", self, o)
    def __radd__(self, o): return self._apply("ADD", o, self)
    def __sub__(self, o): return self._apply("SUB", self, o)
    def __rsub__(self, o): return self._apply("SUB", o, self)
    def __mul__(self, o): return self._apply("MUL", self, o)
    def __rmul__(self, o): return self._apply("MUL", o, self)
    def __truediv__(self, o): return self._apply("DIV", self, o)
    def __rtruediv__(self, o): return self._apply("DIV", o, self)
    def __pow__(self, o): return self._apply("POW", self, o)
    def __rpow__(self, o): return self._apply("POW", o, self)
    def __neg__(self): return ComplexParticles(-self.re, -self.im)
    def conj(self): return ComplexParticles(self.re, -self.im)
    def __abs__(self): return P.hypot(self.re, self.im)
    def abs2(self): return self.re * self.re + self.im * self.im
    def angle(self): return P.atan(self.im, self.re)


def _c(z) -> ComplexParticles:
    if isinstance(z, ComplexParticles):
        return z
    if isinstance(z, P.Particles):
        return ComplexParticles(z)
    raise L.MethodError(L.MCM_ERR_INVALID, "expected Complex{Particles} or Particles")


def complex_particles(re, im=None) -> ComplexParticles:
    """`complex(re, im)` for Particles arguments."""
    return ComplexParticles(re, im)


def csqrt(z): z = _c(z); return z._apply("SQRT", z)
def cexp(z): z = _c(z); return z._apply("EXP", z)
def cexp10(z): z = _c(z); return z._apply("EXP10", z)
def clog(z): z = _c(z); return z._apply("LOG", z)
def clog10(z): z = _c(z); return z._apply("LOG10", z)
def csin(z): z = _c(z); return z._apply("SIN", z)
def ccos(z): z = _c(z); return z._apply("COS", z)
def ctan(z): z = _c(z); return z._apply("TAN", z)


# statistics act on the two components (complex.jl:96-98)
def pmean(z): return complex(P.pmean(z.re), P.pmean(z.im))
def pstd(z): return complex(P.pstd(z.re), P.pstd(z.im))
def pmaximum(z): return complex(P.pmaximum(z.re), P.pmaximum(z.im))
def pminimum(z): return complex(P.pminimum(z.re), P.pminimum(z.im))
def psum(z): return complex(P.psum(z.re), P.psum(z.im))
def isfinite(z): return bool(np.isfinite(P.pmean(z.re)) and np.isfinite(P.pmean(z.im)))      # complex.jl:83 with particles.jl:345
def isinf(z): return bool(np.isinf(P.pmean(z.re)) or np.isinf(P.pmean(z.im)))                 # complex.jl:82
