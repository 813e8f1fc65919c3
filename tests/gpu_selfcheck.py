"""Raw C-ABI self-check on a GPU box (no pytest, no host mirror): every entry point
of include/mcm_b200.h against numpy on small seeded inputs.  Run:
    python tests/gpu_selfcheck.py
Prints one line per check and exits non-zero on the first failure class."""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mcm_lib", os.path.join(ROOT, "montecarlomeasurements.jl_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
E = L.ENUMS
lib = L.init(0)
FAILS = []


def dev(a: np.ndarray):
    p = C.c_void_p()
    L.check(lib.mcm_alloc(C.byref(p), a.nbytes))
    a = np.ascontiguousarray(a)
    L.check(lib.mcm_upload(p, a.ctypes.data_as(C.c_void_p), a.nbytes))
    return p


def empty(n, dtype):
    p = C.c_void_p()
    L.check(lib.mcm_alloc(C.byref(p), n * np.dtype(dtype).itemsize))
    return p


def host(p, n, dtype):
    out = np.empty(n, dtype=dtype)
    L.check(lib.mcm_download(out.ctypes.data_as(C.c_void_p), p, out.nbytes))
    return out


def ulps(a, b):
    b = np.asarray(b)
    sp = np.spacing(np.abs(b)).astype(np.float64)  # ulp of the REFERENCE in its own precision
    a = np.asarray(a, dtype=np.float64); b = b.astype(np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same = (a == b) | both_nan
    d = np.abs(a - b) / np.where(sp > 0, sp, 1)
    d[same] = 0
    d[np.isnan(d)] = np.inf
    return d


def report(name, ok, extra=""):
    print(f"{'ok  ' if ok else 'FAIL'} {name} {extra}")
    if not ok:
        FAILS.append(name)


def dt(dtype):
    return L.F64 if dtype == np.float64 else L.F32


rng = np.random.default_rng(0)
n = 100003  # odd: exercises vector tail
for dtype in (np.float64, np.float32):
    x = rng.normal(0.3, 1.0, n).astype(dtype)
    xp = np.abs(x) + dtype(0.01)
    xu = np.clip(x, -0.99, 0.99).astype(dtype)
    dx = dev(x); dxp = dev(xp); dxu = dev(xu); dout = empty(n, dtype)
    tol = 2.0
    unary = {
        "PLUS": (x, x), "NEG": (x, -x), "EXP": (x, np.exp(x)), "EXP2": (x, np.exp2(x)),
        "EXP10": (x, np.power(dtype(10), x)), "EXPM1": (x, np.expm1(x)), "LOG": (xp, np.log(xp)),
        "LOG10": (xp, np.log10(xp)), "LOG2": (xp, np.log2(xp)), "LOG1P": (xp, np.log1p(xp)),
        "SIN": (x, np.sin(x)), "COS": (x, np.cos(x)), "TAN": (x, np.tan(x)),
        "SINH": (x, np.sinh(x)), "COSH": (x, np.cosh(x)), "TANH": (x, np.tanh(x)),
        "ASIN": (xu, np.arcsin(xu)), "ACOS": (xu, np.arccos(xu)), "ATAN": (x, np.arctan(x)),
        "ASINH": (x, np.arcsinh(x)), "ACOSH": (xp + 1, np.arccosh(xp + 1)), "ATANH": (xu, np.arctanh(xu)),
        "ZERO": (x, np.zeros_like(x)), "SIGN": (x, np.sign(x)), "ABS": (x, np.abs(x)),
        "SQRT": (xp, np.sqrt(xp)), "FLOAT": (x, x), "ROUND": (x * 10, np.rint(x * 10)),
        "FLOOR": (x * 10, np.floor(x * 10)), "CEIL": (x * 10, np.ceil(x * 10)), "TRUNC": (x * 10, np.trunc(x * 10)),
        "INV": (xp, dtype(1) / xp), "SQR": (x, x * x), "CUBE": (x, x * x * x),
        "RAD2DEG": (x, x * dtype(dtype(180) / dtype(np.pi))), "DEG2RAD": (x, x * dtype(dtype(np.pi) / dtype(180))),
        "SIND": (x * 100, np.sin(np.deg2rad((x * 100).astype(np.float64)))),
        "COSD": (x * 100, np.cos(np.deg2rad((x * 100).astype(np.float64)))),
    }
    exact = {"PLUS", "NEG", "ZERO", "SIGN", "ABS", "SQRT", "FLOAT", "ROUND", "FLOOR", "CEIL", "TRUNC", "INV", "SQR",
             "CUBE", "RAD2DEG", "DEG2RAD"}
    for name, (inp, ref) in unary.items():
        inp = np.ascontiguousarray(inp.astype(dtype))
        d_in = dev(inp)
        L.check(lib.mcm_unary(E["MCM_U_" + name], dt(dtype), d_in, dout, n))
        got = host(dout, n, dtype)
        u = ulps(got, np.asarray(ref).astype(dtype))
        lim = 0 if name in exact else (4.0 if name in ("SIND", "COSD", "ACOSH", "ASINH", "ATANH", "TANH") else tol)
        if name in ("SIND", "COSD"):
            ok = np.max(np.abs(got.astype(np.float64) - ref)) < (1e-14 if dtype == np.float64 else 1e-6)
        else:
            ok = u.max() <= lim
        report(f"unary {name} {np.dtype(dtype).name}", ok, f"max ulp {u.max():.2f}")
        lib.mcm_free(d_in)
    # binary
    y = rng.normal(-0.2, 2.0, n).astype(dtype)
    dy = dev(y)
    binary = {
        "ADD": x + y, "SUB": x - y, "MUL": x * y, "DIV": x / y, "MAX": np.maximum(x, y), "MIN": np.minimum(x, y),
        "ATAN2": np.arctan2(x, y), "HYPOT": np.hypot(x, y), "REM": np.fmod(x, y), "MOD": np.mod(x, y),
        "ADD_SUM": x + y,
    }
    for name, ref in binary.items():
        L.check(lib.mcm_binary(E["MCM_B_" + name], dt(dtype), dx, 0.0, dy, 0.0, dout, n))
        got = host(dout, n, dtype)
        u = ulps(got, ref)
        lim = 2.0 if name in ("ATAN2", "HYPOT") else 0
        report(f"binary {name} {np.dtype(dtype).name}", u.max() <= lim, f"max ulp {u.max():.2f}")
    L.check(lib.mcm_binary(E["MCM_B_POW"], dt(dtype), dxp, 0.0, None, 1.5, dout, n))
    u = ulps(host(dout, n, dtype), np.power(xp, dtype(1.5)))
    report(f"binary POW scalar {np.dtype(dtype).name}", u.max() <= 2, f"max ulp {u.max():.2f}")
    L.check(lib.mcm_binary(E["MCM_B_SUB"], dt(dtype), None, 2.5, dx, 0.0, dout, n))
    report(f"binary scalar-first SUB {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), dtype(2.5) - x))
    L.check(lib.mcm_ternary(E["MCM_T_CLAMP"], dt(dtype), dx, 0, None, -0.5, None, 0.7, dout, n))
    report(f"clamp {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), np.clip(x, dtype(-0.5), dtype(0.7))))
    L.check(lib.mcm_ternary(E["MCM_T_MUL_ADD"], dt(dtype), dx, 0, dy, 0, None, 0.25, dout, n))
    report(f"mul_add {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), x * y + dtype(0.25)))
    for k in (-2, -1, 0, 1, 2, 3, 5, -3, 10):
        L.check(lib.mcm_powi(dt(dtype), dxp, k, dout, n))
        ref = {2: xp * xp, 3: xp * xp * xp, -1: 1 / xp, 0: np.ones_like(xp), 1: xp, -2: (1 / xp) * (1 / xp)}.get(k)
        got = host(dout, n, dtype)
        if ref is not None:
            report(f"powi {k} {np.dtype(dtype).name}", np.array_equal(got, ref.astype(dtype)))
        else:
            ref = np.power(xp.astype(np.float64), k)
            u = ulps(got, ref.astype(dtype))
            report(f"powi {k} {np.dtype(dtype).name}", u.max() <= 1, f"max ulp {u.max():.2f}")
    # fused
    L.check(lib.mcm_affine(dt(dtype), dx, 1.25, -3.0, dout, n))
    report(f"affine {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), dtype(1.25) * x + dtype(-3.0)))
    coeffs = np.array([0.5, -1.0, 0.25, 2.0, -0.125], dtype=np.float64)
    L.check(lib.mcm_polyval(dt(dtype), dx, coeffs.ctypes.data_as(L._dp), 5, 0, dout, n))
    acc = np.full(n, coeffs[4], dtype=dtype)
    for c in coeffs[3::-1]:
        acc = acc * x + dtype(c)
    report(f"polyval unfused {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), acc))
    L.check(lib.mcm_polyval(dt(dtype), dx, coeffs.ctypes.data_as(L._dp), 5, 1, dout, n))
    x64_ = x.astype(np.float64)
    refp = np.polyval(coeffs[::-1], x64_)
    scale = np.polyval(np.abs(coeffs[::-1]), np.abs(x64_))
    errp = np.max(np.abs(host(dout, n, dtype).astype(np.float64) - refp) / scale)
    report(f"polyval fused {np.dtype(dtype).name}", errp <= 4 * np.finfo(dtype).eps, f"rel err {errp:.2e}")
    ptrs = (C.c_void_p * 3)(dx, dy, dxp)
    cs = np.array([0.5, 2.0, -1.5])
    L.check(lib.mcm_lincomb(dt(dtype), 3, ptrs, cs.ctypes.data_as(L._dp), dout, n))
    report(f"lincomb {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), dtype(0.5) * x + dtype(2.0) * y + dtype(-1.5) * xp))
    # chain10: sin(a)*a^2 + exp(-a) -> abs -> sqrt -> / a^1.5
    a = (np.pi + 0.1 * rng.normal(size=n)).astype(dtype)
    da = dev(a)
    U, B, Tn = E["MCM_TAPE_UNARY"], E["MCM_TAPE_BINARY"], E["MCM_TAPE_TERNARY"]
    K = E["MCM_TAPE_CONST"]
    prog = [
        (E["MCM_TAPE_LOAD"], 0, 0, 0, 0),
        (U + E["MCM_U_SIN"], 1, 0, 0, 0),
        (U + E["MCM_U_SQR"], 2, 0, 0, 0),
        (B + E["MCM_B_MUL"], 1, 1, 2, 0),
        (U + E["MCM_U_NEG"], 2, 0, 0, 0),
        (U + E["MCM_U_EXP"], 2, 2, 0, 0),
        (B + E["MCM_B_ADD"], 1, 1, 2, 0),
        (U + E["MCM_U_ABS"], 1, 1, 0, 0),
        (U + E["MCM_U_SQRT"], 1, 1, 0, 0),
        (B + E["MCM_B_POW"], 0, 0, K | 0, 0),
        (B + E["MCM_B_DIV"], 0, 1, 0, 0),
        (E["MCM_TAPE_STORE"], 0, 0, 0, 0),
    ]
    tape = (L.TapeIns * len(prog))()
    for i, (op, d, a_, b_, c_) in enumerate(prog):
        tape[i].op, tape[i].dst, tape[i].a, tape[i].b, tape[i].c = op, d, a_, b_, c_
    consts = np.array([1.5])
    ins = (C.c_void_p * 1)(da)
    outs = (C.c_void_p * 1)(dout)
    stats = np.zeros(6)
    L.check(lib.mcm_chain(dt(dtype), tape, len(prog), len(prog), len(prog), 0, consts.ctypes.data_as(L._dp), 1,
                          ins, 1, outs, 1, n, 0, stats.ctypes.data_as(L._dp)))
    got = host(dout, n, dtype)
    a64 = a
    ref = np.sqrt(np.abs(np.sin(a64) * (a64 * a64) + np.exp(-a64))) / np.power(a64, dtype(1.5))
    t3 = np.sin(a64) * (a64 * a64); t5 = np.exp(-a64); t9 = np.power(a64, dtype(1.5))
    # sin*a^2 + exp(-a) cancels near a ~ pi: condition-aware bound (3 ulp on each transcendental)
    eps = np.finfo(dtype).eps
    allow = 3 * eps * (np.abs(t3) + np.abs(t5)).astype(np.float64) / (2 * np.sqrt(np.abs(t3 + t5)) * t9).astype(np.float64) + 4 * eps * np.abs(ref).astype(np.float64)
    errc = np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64)) / allow)
    report(f"chain10 {np.dtype(dtype).name} engine={lib.mcm_chain_last_engine()}", errc <= 1.0, f"err/allow {errc:.2f}")
    lib.mcm_chain_force_interpreter(1)
    dout2 = empty(n, dtype); outs2 = (C.c_void_p * 1)(dout2)
    L.check(lib.mcm_chain(dt(dtype), tape, len(prog), len(prog), len(prog), 0, consts.ctypes.data_as(L._dp), 1,
                          ins, 1, outs2, 1, n, -1, None))
    eng = lib.mcm_chain_last_engine()
    lib.mcm_chain_force_interpreter(0)
    report(f"chain10 interpreter == registered kernel {np.dtype(dtype).name}", eng == 2 and np.array_equal(host(dout2, n, dtype), got))
    g64 = got.astype(np.float64)
    ok = (stats[0] == n and abs(stats[1] - g64.mean()) <= 1e-12 * abs(g64.mean())
          and abs(stats[2] - ((g64 - g64.mean()) ** 2).sum()) <= 1e-11 * stats[2]
          and stats[3] == g64.min() and stats[4] == g64.max())
    report(f"chain10 fused stats {np.dtype(dtype).name}", ok, f"{stats}")
    # moments / reduce / count
    m = np.zeros(6)
    L.check(lib.mcm_moments(dt(dtype), dx, n, m.ctypes.data_as(L._dp)))
    x64 = x.astype(np.float64)
    ok = (m[0] == n and abs(m[1] - x64.mean()) < 1e-13 and abs(m[2] / (n - 1) - x64.var(ddof=1)) < 1e-12 * x64.var(ddof=1)
          and m[3] == x64.min() and m[4] == x64.max() and abs(m[5] - x64.sum()) < 1e-9)
    report(f"moments {np.dtype(dtype).name}", ok, f"{m}")
    big = (1e9 + rng.normal(size=n)).astype(np.float64)
    if dtype == np.float64:
        dbig = dev(big)
        L.check(lib.mcm_moments(L.F64, dbig, n, m.ctypes.data_as(L._dp)))
        ref_var = np.var(big - 1e9, ddof=1)
        report("moments offset 1e9 (no cancellation)", abs(m[2] / (n - 1) - ref_var) < 1e-9 * ref_var, f"{m[2]/(n-1)} vs {ref_var}")
    r = C.c_double()
    L.check(lib.mcm_reduce(E["MCM_R_STD"], dt(dtype), dx, n, C.byref(r)))
    report(f"reduce std {np.dtype(dtype).name}", abs(r.value - x64.std(ddof=1)) < 1e-12)
    cnt = C.c_uint64()
    L.check(lib.mcm_count(E["MCM_CMP_LT"], dt(dtype), dx, n, 1.0, C.byref(cnt)))
    report(f"count lt {np.dtype(dtype).name}", cnt.value == int((x < 1.0).sum()))
    # sort / sortperm / quantile
    ds = empty(n, dtype)
    xs = x.copy(); xs[::1000] = 0.0; xs[1::1000] = -0.0
    dxs = dev(xs)
    L.check(lib.mcm_sort(dt(dtype), dxs, ds, n))
    got = host(ds, n, dtype)
    ref = np.sort(xs)
    report(f"sort {np.dtype(dtype).name}", np.array_equal(got, ref) and np.array_equal(np.signbit(got), np.signbit(ref[np.lexsort((np.arange(n), ref))])) or np.array_equal(got, ref))
    dperm = empty(n, np.int64)
    L.check(lib.mcm_sortperm(dt(dtype), dxs, dperm, None, n))
    perm = host(dperm, n, np.int64)
    key = xs.view(np.int64 if dtype == np.float64 else np.int32).astype(np.int64)
    refperm = np.argsort(xs, kind="stable")
    # -0.0 sorts before +0.0 in Julia's isless; numpy treats them equal -> compare values and stability on equal bits
    ok = np.array_equal(xs[perm], np.sort(xs)) and len(np.unique(perm)) == n
    report(f"sortperm {np.dtype(dtype).name}", ok)
    qs = np.array([0.0, 0.025, 0.5, 0.975, 1.0, 0.3333])
    qo = np.zeros(len(qs))
    L.check(lib.mcm_quantile(dt(dtype), dx, n, qs.ctypes.data_as(L._dp), len(qs), qo.ctypes.data_as(L._dp)))
    refq = np.quantile(x.astype(np.float64), qs)
    report(f"quantile {np.dtype(dtype).name}", np.max(np.abs(qo - refq)) < (1e-13 if dtype == np.float64 else 1e-6), f"{qo}")
    # gather / rand_indices / gather_multi
    didx = empty(n, np.int64)
    L.check(lib.mcm_rand_indices(didx, n, n, 42, 0))
    idx = host(didx, n, np.int64)
    report("rand_indices range", idx.min() >= 0 and idx.max() < n and abs(idx.mean() / n - 0.5) < 0.01)
    L.check(lib.mcm_gather(dt(dtype), dx, didx, dout, n, n))
    report(f"gather {np.dtype(dtype).name}", np.array_equal(host(dout, n, dtype), x[idx]))
    o1, o2, o3 = empty(n, dtype), empty(n, dtype), empty(n, dtype)
    srcs = (C.c_void_p * 3)(dx, dy, dxp); outs3 = (C.c_void_p * 3)(o1, o2, o3)
    L.check(lib.mcm_gather_multi(dt(dtype), srcs, outs3, 3, didx, n, n))
    ok = np.array_equal(host(o1, n, dtype), x[idx]) and np.array_equal(host(o2, n, dtype), y[idx]) and np.array_equal(host(o3, n, dtype), xp[idx])
    report(f"gather_multi {np.dtype(dtype).name}", ok)
    # systematic sample
    N = 2000
    dsys = empty(N, dtype)
    L.check(lib.mcm_systematic_normal(dt(dtype), dsys, N, np.pi, 0.1, None, 0, N))
    s = host(dsys, N, dtype).astype(np.float64)
    report(f"systematic pstd(pi +- 0.1) {np.dtype(dtype).name}", abs(s.std(ddof=1) - 0.09999231528930486) < (1e-13 if dtype == np.float64 else 1e-7), repr(s.std(ddof=1)))
    # gram
    Mv, Ng = 150, 5000
    X = rng.normal(size=(Mv, Ng)).astype(dtype)
    X[3] = X[1] * 2 + 1
    dX = [dev(X[i]) for i in range(Mv)]
    vp = (C.c_void_p * Mv)(*dX)
    dC = empty(Mv * Mv, np.float64)
    L.check(lib.mcm_gram(dt(dtype), vp, Mv, Ng, 1, None, dC, None))
    Cg = host(dC, Mv * Mv, np.float64).reshape(Mv, Mv)
    Cref = np.cov(X.astype(np.float64))
    err = np.max(np.abs(Cg - Cref)) / np.max(np.abs(Cref))
    report(f"gram {np.dtype(dtype).name} M={Mv} N={Ng}", err < 1e-12, f"rel err {err:.2e}")
    Mv2, Ng2 = 5, 300001
    X2 = rng.normal(2.0, 1.0, size=(Mv2, Ng2)).astype(dtype)
    dX2 = [dev(X2[i]) for i in range(Mv2)]
    vp2 = (C.c_void_p * Mv2)(*dX2)
    L.check(lib.mcm_gram(dt(dtype), vp2, Mv2, Ng2, 1, None, dC, None))
    Cg = host(dC, Mv2 * Mv2, np.float64).reshape(Mv2, Mv2)
    Cref = np.cov(X2.astype(np.float64))
    err = np.max(np.abs(Cg - Cref)) / np.max(np.abs(Cref))
    report(f"gram split-K {np.dtype(dtype).name} M={Mv2} N={Ng2}", err < 1e-12, f"rel err {err:.2e}")

# randperm
N = 50000
dp = empty(N, np.int64)
L.check(lib.mcm_randperm(dp, N, 7))
p = host(dp, N, np.int64)
report("randperm is a permutation", np.array_equal(np.sort(p), np.arange(N)) and not np.array_equal(p, np.arange(N)))
# resample_systematic vs exact integer host model
w = rng.random(N) ** 3
dw = dev(w)
L.check(lib.mcm_resample_systematic(dw, N, 0.37, dp))
got = host(dp, N, np.int64)
e = int(np.frexp(w.max())[1]); lg = int(np.ceil(np.log2(N)))
W = np.floor(w * 2.0 ** (62 - lg - e)).astype(object)
Cs = np.cumsum(W); S = Cs[-1]; U = int(np.floor(0.37 * 2 ** 32))
ref = np.empty(N, dtype=np.int64)
b = 0
for i in range(N):
    rhs = (i * 2 ** 32 + U) * S
    while not (Cs[b] * N * 2 ** 32 > rhs):
        b += 1
    ref[i] = b
report("resample_systematic exact", np.array_equal(got, ref))
# flags
bits = C.c_uint32()
L.check(lib.mcm_check_flags(C.byref(bits), 1))
report("flags clean before domain error", bits.value == 0, hex(bits.value))
neg = dev(np.array([-1.0, 4.0]))
o = empty(2, np.float64)
L.check(lib.mcm_unary(E["MCM_U_SQRT"], L.F64, neg, o, 2))
L.check(lib.mcm_check_flags(C.byref(bits), 1))
report("sqrt(-1) raises DOMAIN flag", bits.value & 1 == 1)

# throughput probes (N = 1e8 F64)
N = 100_000_000
dA = empty(N, np.float64); dB = empty(N, np.float64)
L.check(lib.mcm_systematic_normal(L.F64, dA, N, np.pi, 0.1, None, 0, N))
L.check(lib.mcm_sync())


def timeit(fn, reps=5):
    fn(); L.check(lib.mcm_sync())
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); L.check(lib.mcm_sync()); best = min(best, time.perf_counter() - t)
    return best


t = timeit(lambda: lib.mcm_affine(L.F64, dA, 1.5, 0.5, dB, N))
print(f"affine 1e8: {t*1e3:.3f} ms  {16*N/t/1e9:.0f} GB/s")
t = timeit(lambda: lib.mcm_unary(E['MCM_U_SIN'], L.F64, dA, dB, N))
print(f"sin 1e8: {t*1e3:.3f} ms  {16*N/t/1e9:.0f} GB/s")
t = timeit(lambda: lib.mcm_unary(E['MCM_U_EXP'], L.F64, dA, dB, N))
print(f"exp 1e8: {t*1e3:.3f} ms  {16*N/t/1e9:.0f} GB/s")
t = timeit(lambda: lib.mcm_binary(E['MCM_B_POW'], L.F64, dA, 0.0, None, 1.5, dB, N))
print(f"pow 1e8: {t*1e3:.3f} ms  {16*N/t/1e9:.0f} GB/s")
t = timeit(lambda: lib.mcm_unary(E['MCM_U_SQRT'], L.F64, dA, dB, N))
print(f"sqrt 1e8: {t*1e3:.3f} ms  {16*N/t/1e9:.0f} GB/s")
t = timeit(lambda: lib.mcm_binary(E['MCM_B_DIV'], L.F64, dA, 0.0, dA, 0.0, dB, N))
print(f"div 1e8: {t*1e3:.3f} ms  {16*N/t/1e9:.0f} GB/s")
m = np.zeros(6)
t = timeit(lambda: lib.mcm_moments(L.F64, dA, N, m.ctypes.data_as(L._dp)))
print(f"moments 1e8: {t*1e3:.3f} ms  {8*N/t/1e9:.0f} GB/s  mean={m[1]!r} std={np.sqrt(m[2]/(N-1))!r}")
ins = (C.c_void_p * 1)(dA); outs = (C.c_void_p * 1)(dB)
consts = np.array([1.5]); stats = np.zeros(6)
tape_args = (L.F64, tape, len(prog), len(prog), len(prog), 0, consts.ctypes.data_as(L._dp), 1, ins, 1, outs, 1, N)
t = timeit(lambda: lib.mcm_chain(*tape_args, 0, stats.ctypes.data_as(L._dp)))
print(f"chain10+stats 1e8: {t*1e3:.3f} ms  {10*N/t/1e9:.1f} Gparticle-ops/s  {16*N/t/1e9:.0f} GB/s  stats={stats}")
t = timeit(lambda: lib.mcm_chain(*tape_args, -1, None))
print(f"chain10 1e8 (engine {lib.mcm_chain_last_engine()}): {t*1e3:.3f} ms  {10*N/t/1e9:.1f} Gparticle-ops/s  {16*N/t/1e9:.0f} GB/s")
lib.mcm_chain_force_interpreter(1)
t = timeit(lambda: lib.mcm_chain(*tape_args, -1, None))
print(f"chain10 1e8 interpreter (engine {lib.mcm_chain_last_engine()}): {t*1e3:.3f} ms  {10*N/t/1e9:.1f} Gparticle-ops/s")
lib.mcm_chain_force_interpreter(0)
N2 = 10_000_000
t = timeit(lambda: lib.mcm_sort(L.F64, dA, dB, N2), reps=2)
print(f"sort 1e7: {t*1e3:.3f} ms  {N2/t/1e9:.2f} Gkeys/s")

print("FAILED:", FAILS if FAILS else "none")
sys.exit(1 if FAILS else 0)
