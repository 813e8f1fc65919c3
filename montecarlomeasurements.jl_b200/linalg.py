"""Per-particle small-matrix operations on uncertain matrices.

Reference: `ℝⁿ2ℝⁿ_function` / `ℝⁿ2ℂⁿ_function` (src/particles.jl:241-282) and the methods built on them —
`exp(::Matrix{<:AbstractParticles})`, `exp!`, `lyap` (:603-607), `eigvals` (:395-406), plus `\\`, `inv`, `det`,
`svdvals` which the reference reaches through generic LinearAlgebra code or `bymap`.  The reference loops over the
particle index on the host and calls LAPACK N times; here one device call (`mcm_batched_linalg`) handles all N
matrices, one thread per particle.

An uncertain matrix is a nested list / tuple / object ndarray whose entries are `Particles` or plain numbers (numbers
become constant particles, as `promote` does in the reference).  Results come back as nested lists of `Particles`."""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from . import _lib as L
from . import particles as P

E = L.ENUMS
MAX_N = 16


def _as_table(M, what: str):
    """-> (rows, cols, flat list of Particles, N, dtype); numbers are promoted to constant particles."""
    if isinstance(M, np.ndarray):
        arr = M.astype(object)
    else:
        rows = [list(r) if isinstance(r, (list, tuple, np.ndarray)) else [r] for r in M]
        if len({len(r) for r in rows}) != 1:
            raise L.ArgumentError(L.MCM_ERR_INVALID, f"{what}: ragged matrix")
        arr = np.empty((len(rows), len(rows[0])), dtype=object)
        for i, r in enumerate(rows):
            for j, v in enumerate(r):
                arr[i, j] = v
    if arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    parts = [v for v in arr.ravel() if isinstance(v, P.Particles)]
    if not parts:
        raise L.MethodError(L.MCM_ERR_INVALID, f"{what}: no Particles among the entries (use numpy for plain matrices)")
    N, dt = parts[0].N, parts[0].dtype
    flat = []
    for v in arr.ravel():
        if isinstance(v, P.Particles):
            if v.N != N:
                raise L.MethodError(L.MCM_ERR_INVALID, f"{what}: all entries need the same number of particles")
            flat.append(v)
        elif isinstance(v, numbers.Real):
            flat.append(P.fill(float(v), N, P._NP[dt]))
        else:
            raise L.MethodError(L.MCM_ERR_INVALID, f"{what}: unsupported entry type {type(v).__name__}")
    if any(v.dtype != flat[0].dtype for v in flat):
        raise L.MethodError(L.MCM_ERR_INVALID, f"{what}: mixed element types")
    return arr.shape[0], arr.shape[1], flat, N, flat[0].dtype


def _ptrs(ps):
    return (C.c_void_p * len(ps))(*[C.c_void_p(p.ptr) for p in ps])


def _call(op: str, A, B=None, n_out=None, im=False, who="linalg"):
    r, c, fa, N, dt = _as_table(A, who)
    if r != c:
        raise L.ArgumentError(L.MCM_ERR_INVALID, f"{who}: matrix must be square (DimensionMismatch in the reference)")
    if r > MAX_N:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, f"{who}: device path handles n <= {MAX_N}; use bymap for larger matrices")
    m, fb = 0, None
    if B is not None:
        rb, m, fb, Nb, dtb = _as_table(B, who)
        if rb != r or Nb != N or dtb != dt:
            raise L.ArgumentError(L.MCM_ERR_INVALID, f"{who}: right-hand side does not match the matrix")
    k = n_out(r, m)
    outs = [P._empty(N, dt) for _ in range(k)]
    outs_im = [P._empty(N, dt) for _ in range(r)] if im else None
    L.check(L.lib().mcm_batched_linalg(E["MCM_LA_" + op], dt, r, m, _ptrs(fa), _ptrs(fb) if fb else None, _ptrs(outs),
                                       _ptrs(outs_im) if im else None, N))
    return r, m, outs, outs_im


def _grid(flat, rows, cols):
    return [[flat[i * cols + j] for j in range(cols)] for i in range(rows)]


def solve(A, B):
    """`A \\ B` per particle (LU with partial pivoting).  B: vector (list of n entries) or n x m matrix."""
    vec = not isinstance(B[0], (list, tuple, np.ndarray)) if not isinstance(B, np.ndarray) else B.ndim == 1
    r, m, outs, _ = _call("SOLVE", A, [[b] for b in B] if vec else B, lambda n, m: n * m, who="solve")
    return outs if vec else _grid(outs, r, m)


def inv(A):
    r, _, outs, _ = _call("INV", A, None, lambda n, m: n * n, who="inv")
    return _grid(outs, r, r)


def det(A) -> P.Particles:
    return _call("DET", A, None, lambda n, m: 1, who="det")[2][0]


def expm(A):
    """`exp(A)` / `LinearAlgebra.exp!(A)` for a matrix of particles (particles.jl:605-606)."""
    r, _, outs, _ = _call("EXP", A, None, lambda n, m: n * n, who="exp")
    return _grid(outs, r, r)


def eigvals(A):
    """`eigvals(A)` (particles.jl:395-406): always complex in the reference; returned here as (real, imag) lists of
    Particles, each particle's spectrum ordered by (real, imaginary part)."""
    _, _, outs, outs_im = _call("EIGVALS", A, None, lambda n, m: n, im=True, who="eigvals")
    return outs, outs_im


def eigvals_sym(A):
    """Eigenvalues of a symmetric uncertain matrix (what `eigvals(Symmetric(A))` gives per particle), ascending."""
    return _call("EIGVALS_SYM", A, None, lambda n, m: n, who="eigvals_sym")[2]


def lyap(A, Cm):
    """`lyap(A, C)` (particles.jl:607): X with A X + X A' + C = 0, per particle (n <= 8: the Kronecker system of order n^2 is factored per thread)."""
    r, _, outs, _ = _call("LYAP", A, Cm, lambda n, m: n * n, who="lyap")
    return _grid(outs, r, r)


def svdvals(A):
    return _call("SVDVALS", A, None, lambda n, m: n, who="svdvals")[2]


# ---- matrices with Complex{Particles} entries; log of a real uncertain matrix (complex.jl:104-181, particles.jl:603) ----
def _as_complex_table(M, who: str):
    """-> (n, re entries, im entries or None, N, dtype).  Entries: ComplexParticles, Particles, complex or real numbers."""
    from .complexp import ComplexParticles
    if isinstance(M, np.ndarray):
        arr = M.astype(object)
    else:                        # element by element: np.array() would download Particles through __array__
        rows = [list(r) if isinstance(r, (list, tuple, np.ndarray)) else [r] for r in M]
        if len({len(r) for r in rows}) != 1:
            raise L.ArgumentError(L.MCM_ERR_INVALID, f"{who}: ragged matrix")
        arr = np.empty((len(rows), len(rows[0])), dtype=object)
        for i, r in enumerate(rows):
            for j, v in enumerate(r):
                arr[i, j] = v
    if arr.ndim != 2 or arr.shape[0] != arr.shape[1]:
        raise L.ArgumentError(L.MCM_ERR_INVALID, f"{who}: matrix must be square (DimensionMismatch in the reference)")
    n = arr.shape[0]
    if n > MAX_N:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, f"{who}: device path handles n <= {MAX_N}; use bymap for larger matrices")
    ref = None
    for v in arr.ravel():
        if isinstance(v, ComplexParticles):
            ref = v.re
        elif isinstance(v, P.Particles):
            ref = v
        if ref is not None:
            break
    if ref is None:
        raise L.MethodError(L.MCM_ERR_INVALID, f"{who}: no Particles among the entries (use numpy for plain matrices)")
    N, dt = ref.N, ref.dtype
    any_complex = any(isinstance(v, (ComplexParticles, complex)) and not isinstance(v, numbers.Real) for v in arr.ravel())
    re, im = [], []
    for v in arr.ravel():
        if isinstance(v, ComplexParticles):
            a, b = v.re, v.im
        elif isinstance(v, P.Particles):
            a, b = v, None
        elif isinstance(v, numbers.Real):
            a, b = P.fill(float(v), N, P._NP[dt]), None
        elif isinstance(v, numbers.Complex):
            a, b = P.fill(float(v.real), N, P._NP[dt]), P.fill(float(v.imag), N, P._NP[dt])
        else:
            raise L.MethodError(L.MCM_ERR_INVALID, f"{who}: unsupported entry type {type(v).__name__}")
        if a.N != N or a.dtype != dt or (b is not None and (b.N != N or b.dtype != dt)):
            raise L.MethodError(L.MCM_ERR_INVALID, f"{who}: all entries need the same number of particles and element type")
        re.append(a)
        im.append(b if b is not None else (P.fill(0.0, N, P._NP[dt]) if any_complex else None))
    return n, re, (im if any_complex else None), N, dt


def _ccall(op: str, A, nout, who: str, real_out: bool = False):
    from .complexp import ComplexParticles
    n, re, im, N, dt = _as_complex_table(A, who)
    k = nout(n)
    ore = P._empty_many(k, N, dt)
    oim = None if real_out else P._empty_many(k, N, dt)
    L.check(L.lib().mcm_batched_linalg_complex(E["MCM_LAC_" + op], dt, n, _ptrs(re), _ptrs(im) if im is not None else None,
                                               _ptrs(ore), _ptrs(oim) if oim is not None else None, N))
    if real_out:
        return n, ore
    return n, [ComplexParticles(a, b) for a, b in zip(ore, oim)]


def cexpm(A):
    """`exp(A)` / `exp!(A)` for a Matrix{Complex{Particles}} (complex.jl:129-130): n x n nested list of ComplexParticles."""
    n, out = _ccall("EXP", A, lambda n: n * n, "exp")
    return _grid(out, n, n)


def logm(A):
    """`log(A)` for a real uncertain matrix (particles.jl:603, complex-valued like the reference's ℝⁿ2ℂⁿ_function) or a
    Matrix{Complex{Particles}} (complex.jl:131): principal matrix logarithm per particle."""
    n, out = _ccall("LOG", A, lambda n: n * n, "log")
    return _grid(out, n, n)


def cdet(A):
    """`det(A)` for a Matrix{Complex{Particles}} (complex.jl:145): one ComplexParticles."""
    return _ccall("DET", A, lambda n: 1, "det")[1][0]


def ceigvals(A):
    """`eigvals(A)` for a Matrix{Complex{Particles}} (complex.jl:147-163): list of ComplexParticles, each particle's
    spectrum ordered by (real, imaginary part)."""
    return _ccall("EIGVALS", A, lambda n: n, "eigvals")[1]


def csvdvals(A):
    """`svdvals(A)` for a Matrix{Complex{Particles}} (complex.jl:165-181): list of Particles, descending."""
    return _ccall("SVDVALS", A, lambda n: n, "svdvals", real_out=True)[1]
