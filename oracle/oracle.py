"""Python face of the CPU oracle (oracle/mcm_oracle.c) plus the mpmath truth tier.

TEST INFRASTRUCTURE ONLY — see the header of mcm_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_ref", "libmcm_oracle.so")
REFCPU_SO = os.path.join(HERE, "_ref", "libref_cpu.so")
F32, F64 = 0, 1


def build(force: bool = False):
    """Compile the C restatement (and the SLEEF CPU baseline) into oracle/_ref/."""
    if force or not (os.path.exists(ORACLE_SO) and os.path.exists(REFCPU_SO)) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "mcm_oracle.c")) or \
            os.path.getmtime(REFCPU_SO) < os.path.getmtime(os.path.join(HERE, "ref_cpu.cpp")):
        p = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + p.stdout + p.stderr)


def _enums():
    txt = open(os.path.join(HERE, "..", "include", "mcm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", txt, flags=re.S):
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = [s.strip() for s in item.split("=", 1)]
                val = int(expr.rstrip("u"), 0)
            else:
                name, val = item, val + 1
            out[name] = val
    return out


E = _enums()
_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_norminvcdf.restype = C.c_double
        _lib.orc_norminvcdf.argtypes = [C.c_double]
    return _lib


def _dt(a):
    return F64 if a.dtype == np.float64 else F32


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _chk(rc):
    if rc != 0:
        raise ValueError(f"oracle status {rc}")


def unary(name: str, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    _chk(lib().orc_unary(E["MCM_U_" + name], _dt(x), _p(x), _p(out), C.c_size_t(x.size)))
    return out


def binary(name: str, a, b) -> np.ndarray:
    arr = a if isinstance(a, np.ndarray) else b
    A = np.ascontiguousarray(a) if isinstance(a, np.ndarray) else None
    B = np.ascontiguousarray(b) if isinstance(b, np.ndarray) else None
    out = np.empty_like(arr)
    _chk(lib().orc_binary(E["MCM_B_" + name], _dt(arr), _p(A), C.c_double(0.0 if A is not None else float(a)),
                          _p(B), C.c_double(0.0 if B is not None else float(b)), _p(out), C.c_size_t(arr.size)))
    return out


def ternary(name: str, a, b, c) -> np.ndarray:
    arr = next(x for x in (a, b, c) if isinstance(x, np.ndarray))
    args = []
    for x in (a, b, c):
        if isinstance(x, np.ndarray):
            args += [_p(np.ascontiguousarray(x)), C.c_double(0.0)]
        else:
            args += [None, C.c_double(float(x))]
    out = np.empty_like(arr)
    _chk(lib().orc_ternary(E["MCM_T_" + name], _dt(arr), *args, _p(out), C.c_size_t(arr.size)))
    return out


def powi(x: np.ndarray, k: int) -> np.ndarray:
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    _chk(lib().orc_powi(_dt(x), _p(x), C.c_int(k), _p(out), C.c_size_t(x.size)))
    return out


def moments(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    out = np.zeros(6)
    _chk(lib().orc_moments(_dt(x), _p(x), C.c_size_t(x.size), _p(out)))
    return out


def reduce(kind: str, x: np.ndarray) -> float:
    x = np.ascontiguousarray(x)
    r = C.c_double()
    _chk(lib().orc_reduce(E["MCM_R_" + kind], _dt(x), _p(x), C.c_size_t(x.size), C.byref(r)))
    return r.value


def count(cmp: str, x: np.ndarray, thr: float) -> int:
    x = np.ascontiguousarray(x)
    c = C.c_uint64()
    _chk(lib().orc_count(E["MCM_CMP_" + cmp], _dt(x), _p(x), C.c_size_t(x.size), C.c_double(thr), C.byref(c)))
    return int(c.value)


def cov(X: np.ndarray, corrected: bool = True) -> np.ndarray:
    """X: (M, N) array, row j = variable j (i.e. Matrix(v)' of the reference)."""
    X = np.ascontiguousarray(X)
    M, N = X.shape
    ptrs = (C.c_void_p * M)(*[X[j].ctypes.data for j in range(M)])
    out = np.zeros((M, M))
    _chk(lib().orc_cov(_dt(X), ptrs, C.c_int(M), C.c_size_t(N), C.c_int(1 if corrected else 0), _p(out)))
    return out


def sort(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    _chk(lib().orc_sort(_dt(x), _p(x), _p(out), C.c_size_t(x.size)))
    return out


def sortperm(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    out = np.empty(x.size, dtype=np.int64)
    _chk(lib().orc_sortperm(_dt(x), _p(x), _p(out), C.c_size_t(x.size)))
    return out


def quantile(x: np.ndarray, q) -> np.ndarray:
    x = np.ascontiguousarray(x)
    q = np.ascontiguousarray(np.atleast_1d(np.asarray(q, dtype=np.float64)))
    out = np.zeros(q.size)
    _chk(lib().orc_quantile(_dt(x), _p(x), C.c_size_t(x.size), _p(q), C.c_int(q.size), _p(out)))
    return out


def norminvcdf(p) -> np.ndarray:
    p = np.atleast_1d(np.asarray(p, dtype=np.float64))
    f = lib().orc_norminvcdf
    return np.array([f(float(v)) for v in p])


def systematic_normal(N: int, mu: float, sigma: float, dtype=np.float64, perm=None, i0: int = 0, n_total=None,
                      n_local=None) -> np.ndarray:
    n_total = N if n_total is None else n_total
    n = N if n_local is None else n_local
    out = np.empty(n, dtype=dtype)
    pp = np.ascontiguousarray(perm, dtype=np.int64) if perm is not None else None
    _chk(lib().orc_systematic_normal(_dt(out), _p(out), C.c_size_t(n), C.c_double(mu), C.c_double(sigma), _p(pp),
                                     C.c_uint64(i0), C.c_uint64(n_total)))
    return out


def systematic_uniform(N: int, a: float, b: float, dtype=np.float64, perm=None) -> np.ndarray:
    out = np.empty(N, dtype=dtype)
    pp = np.ascontiguousarray(perm, dtype=np.int64) if perm is not None else None
    _chk(lib().orc_systematic_uniform(_dt(out), _p(out), C.c_size_t(N), C.c_double(a), C.c_double(b), _p(pp),
                                      C.c_uint64(0), C.c_uint64(N)))
    return out


def gather(src: np.ndarray, idx: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src); idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty(idx.size, dtype=src.dtype)
    _chk(lib().orc_gather(_dt(src), _p(src), _p(idx), _p(out), C.c_size_t(idx.size), C.c_size_t(src.size)))
    return out


def rand_indices(n: int, N: int, seed: int = 0, offset: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    _chk(lib().orc_rand_indices(_p(out), C.c_size_t(n), C.c_uint64(N), C.c_uint64(seed), C.c_uint64(offset)))
    return out


def randperm(n: int, seed: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    _chk(lib().orc_randperm(_p(out), C.c_size_t(n), C.c_uint64(seed)))
    return out


def index_bijection(n: int, seed: int = 0, column: int = 0, i0: int = 0, n_total=None) -> np.ndarray:
    n_total = n if n_total is None else n_total
    out = np.empty(n, dtype=np.int64)
    lib().orc_index_bijection.restype = None
    lib().orc_index_bijection(_p(out), C.c_size_t(n), C.c_uint64(seed), C.c_uint64(column), C.c_uint64(i0),
                              C.c_uint64(n_total))
    return out


def pgemm(A: np.ndarray, xs: np.ndarray, shift=None, bias=None) -> np.ndarray:
    """ys[j] = bias[j] + sum_k A[j,k] (xs[k] - shift[k]); xs is Mi x n (variable-major); long-double accumulation."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    Mo, Mi = A.shape
    n = xs.shape[1]
    sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    bi = None if bias is None else np.ascontiguousarray(bias, dtype=np.float64)
    ys = np.empty((Mo, n))
    lib().orc_pgemm.restype = None
    lib().orc_pgemm(_p(A), C.c_int(Mo), C.c_int(Mi), _p(xs), _p(sh), _p(bi), _p(ys), C.c_size_t(n))
    return ys


def rand_uniform(n: int, seed: int = 0, offset: int = 0) -> np.ndarray:
    out = np.empty(n)
    lib().orc_rand_uniform.restype = None
    lib().orc_rand_uniform(_p(out), C.c_size_t(n), C.c_uint64(seed), C.c_uint64(offset))
    return out


def autocor(x: np.ndarray, maxlag: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(maxlag)
    lib().orc_autocor.restype = None
    lib().orc_autocor(_p(x), C.c_size_t(x.size), C.c_int(maxlag), _p(out))
    return out


def ess(x: np.ndarray) -> float:
    """src/sampling.jl:26-38 restated."""
    N = len(x)
    ac = autocor(x, min(250, N // 2))
    n = len(ac)
    tau_inv = 1 + 2 * ac[0]
    K = 2
    while K < n - 2:
        d = ac[K - 1] + ac[K]
        if d < 0:
            break
        tau_inv += 2 * d
        K += 2
    return min(1 / tau_inv, 1.0) * N


def resample_systematic(w: np.ndarray, u: float) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty(w.size, dtype=np.int64)
    _chk(lib().orc_resample_systematic(_p(w), C.c_size_t(w.size), C.c_double(u), _p(out)))
    return out


# ---- chains: the reference evaluates one broadcast per primitive (register_primitive.jl:44-58) ----
def chain10(a: np.ndarray) -> np.ndarray:
    """SURVEY.md §8(d) canonical chain, op by op, every temporary rounded to the element type."""
    t1 = unary("SIN", a); t2 = unary("SQR", a); t3 = binary("MUL", t1, t2)
    t4 = unary("NEG", a); t5 = unary("EXP", t4); t6 = binary("ADD", t3, t5)
    t7 = unary("ABS", t6); t8 = unary("SQRT", t7); t9 = binary("POW", a, 1.5)
    return binary("DIV", t8, t9)


def readme_expr(a: np.ndarray) -> np.ndarray:
    return binary("ADD", binary("MUL", unary("SIN", a), unary("SQR", a)), unary("EXP", unary("NEG", a)))


def chain10_bound(a: np.ndarray, ulps_trans: float = 3.0) -> np.ndarray:
    """Absolute error bound for chain10 when each transcendental is within `ulps_trans` ulp: the sum
    sin(a)a^2 + exp(-a) cancels near a ~ pi, which amplifies ulp-level differences."""
    a64 = a.astype(np.float64)
    eps = float(np.finfo(a.dtype).eps)
    t3 = np.sin(a64) * a64 * a64; t5 = np.exp(-a64); t9 = a64 ** 1.5
    y = np.sqrt(np.abs(t3 + t5)) / t9
    return ulps_trans * eps * (np.abs(t3) + np.abs(t5)) / (2 * np.sqrt(np.abs(t3 + t5)) * t9) + (ulps_trans + 2) * eps * np.abs(y)


def pendulum_rk4(th, om, g, Lp, c, dt: float, steps: int):
    """Classical RK4 on θ' = ω, ω' = -(g/L) sin θ - c ω (examples/ode_sim.jl:20-25 + damping), op by op in
    the same association order as montecarlomeasurements.jl_b200/chains.py."""
    th = th.copy(); om = om.copy()
    gL = g / Lp

    def rhs(t, o):
        return o, -(gL * np.sin(t)) - c * o
    for _ in range(steps):
        k1t, k1o = rhs(th, om)
        k2t, k2o = rhs(th + (dt / 2) * k1t, om + (dt / 2) * k1o)
        k3t, k3o = rhs(th + (dt / 2) * k2t, om + (dt / 2) * k2o)
        k4t, k4o = rhs(th + dt * k3t, om + dt * k3o)
        th2 = th + (dt / 6) * (k1t + 2.0 * k2t + 2.0 * k3t + k4t)
        om2 = om + (dt / 6) * (k1o + 2.0 * k2o + 2.0 * k3o + k4o)
        th, om = th2, om2
    return th, om


# ---- mpmath truth tier --------------------------------------------------------------------------
def mp_ulp_error(fn_name: str, x: np.ndarray, y: np.ndarray, y2=None):
    """|y - f(x)| in ulps of the correctly rounded result, f evaluated with mpmath at 100 digits."""
    import mpmath as mp
    mp.mp.dps = 60
    fns = {"sin": mp.sin, "cos": mp.cos, "exp": mp.exp, "log": mp.log, "tan": mp.tan, "asin": mp.asin,
           "acos": mp.acos, "atan": mp.atan, "sinh": mp.sinh, "cosh": mp.cosh, "tanh": mp.tanh,
           "expm1": mp.expm1, "log1p": mp.log1p, "sqrt": mp.sqrt,
           "norminv": lambda p: mp.sqrt(2) * mp.erfinv(2 * p - 1) if p > 0.25 else -mp.sqrt(2) * mp.erfinv(1 - 2 * p)}
    f = fns[fn_name] if fn_name != "pow" else None
    out = np.zeros(len(x))
    for i in range(len(x)):
        t = mp.power(mp.mpf(float(x[i])), mp.mpf(float(y2[i]))) if f is None else f(mp.mpf(float(x[i])))
        tf = float(t)
        ulp = float(np.spacing(np.asarray(abs(tf), dtype=y.dtype))) if tf != 0 else float(np.finfo(y.dtype).tiny)
        out[i] = float(abs(mp.mpf(float(y[i])) - t) / ulp)
    return out


# ---- "reference-shaped" CPU baseline (oracle/ref_cpu.cpp) -----------------------------------------
_ref = None


def refcpu():
    global _ref
    if _ref is None:
        build()
        _ref = C.CDLL(REFCPU_SO)
    return _ref


_ref_bufs = {}


def ref_chain10(a: np.ndarray, threads: int = 1):
    """Run the SLEEF loop-per-primitive baseline once; returns (y, mean, std, seconds).  The result and work vectors
    are allocated (and their pages touched) once per size and reused: the reference pays allocation and page faults
    on fresh vectors for every primitive, this baseline does not charge them."""
    import time
    r = refcpu()
    vl = r.ref_vector_width()
    n = (a.size // vl) * vl
    a = np.ascontiguousarray(a[:n])
    if n not in _ref_bufs:
        _ref_bufs.clear()
        bufs = [np.empty(n) for _ in range(4)]
        for buf in bufs:
            buf.fill(0.0)
        _ref_bufs[n] = bufs
    y, w1, w2, w3 = _ref_bufs[n]
    m = C.c_double(); s = C.c_double()
    t0 = time.perf_counter()
    r.ref_chain10(_p(a), _p(y), _p(w1), _p(w2), _p(w3), C.c_long(n), C.c_int(threads), C.byref(m), C.byref(s))
    dt = time.perf_counter() - t0
    return y, m.value, s.value, dt


# ---- decision layer leftovers (numpy restatements; index/byte work => bit-exact comparisons) ------------------
def union(x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """src/particles.jl:366-368: [p1.particles; p2.particles]."""
    return np.concatenate([x1, x2])


def intersect(x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """src/particles.jl:375-380: common support [max of minima, min of maxima], both filtered in order."""
    mi = max(x1.min(), x2.min())
    ma = min(x1.max(), x2.max())
    keep = lambda x: x[(mi <= x) & (x <= ma)]
    return np.concatenate([keep(x1), keep(x2)])


def julia_argmin(d: np.ndarray) -> int:
    """Base.argmin: first index of the minimum; NaN is smaller than everything (0-based here)."""
    nan = np.isnan(d)
    if nan.any():
        return int(np.argmax(nan))
    return int(np.argmin(d))


def shuffle_and_insert(x: np.ndarray, ind: int, val) -> np.ndarray:
    """src/nominal.jl:1-6 (0-based ind)."""
    y = x.copy()
    y[ind] = y[0]
    y[0] = val
    return y


def with_nominal(x: np.ndarray, val) -> np.ndarray:
    """src/nominal.jl:20-23."""
    return shuffle_and_insert(x, julia_argmin(np.abs(x.astype(np.float64) - val)), val)


def with_nominal_mv(X: np.ndarray, vals: np.ndarray, joint: bool = False):
    """src/nominal.jl:25-29 with X the N x d sample matrix.  Literal form: argmin over the d column sums
    (dims=1); joint=True: argmin over the N row sums (left-to-right accumulation)."""
    D = X.astype(np.float64) - vals[None, :]
    if joint:
        acc = np.zeros(X.shape[0])
        for j in range(X.shape[1]):
            acc = acc + D[:, j] * D[:, j]
        k = julia_argmin(acc)
    else:
        k = julia_argmin((D * D).sum(axis=0))
    return [shuffle_and_insert(X[:, j], k, vals[j]) for j in range(X.shape[1])], k


def rem2pi_truth(x: np.ndarray, mode: str = "RoundNearest") -> np.ndarray:
    """rem2pi(x, r) (Base math.jl; src/particles.jl:359) computed exactly: x - 2pi*round_r(x/2pi) at 1600 bits,
    rounded once to Float64.  Base's own algorithm (rem_pio2_kernel + add22condh on double-doubles) returns this
    value except in near-tie cases; small arguments that Base passes through unchanged are passed through here too."""
    import mpmath as mp
    old = mp.mp.prec
    mp.mp.prec = 1600
    try:
        twopi = 2 * mp.pi
        out = np.empty(len(x))
        for i, v in enumerate(np.asarray(x, dtype=np.float64)):
            if not np.isfinite(v):
                out[i] = np.nan
                continue
            X = mp.mpf(float(v))
            if mode == "RoundNearest":
                r = X if abs(v) < np.pi else X - mp.nint(X / twopi) * twopi
            elif mode == "RoundToZero":
                r = X if abs(v) <= 6.283185307179585 else mp.sign(X) * (abs(X) - mp.floor(abs(X) / twopi) * twopi)
            elif mode == "RoundDown":
                r = X if 0 <= v < 6.283185307179586 else X - mp.floor(X / twopi) * twopi
            elif mode == "RoundUp":
                r = X if -6.283185307179586 < v <= 0 else X - mp.ceil(X / twopi) * twopi
            else:
                raise ValueError(mode)
            out[i] = float(r)
        return out
    finally:
        mp.mp.prec = old
