"""Probe: tcgen05 Ozaki Gram vs the DMMA Gram (accuracy + timing)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
def run(M, n, timing=False, offset=0.0):
    base = [mcm.systematic_sample(n, mcm.Normal(offset, 1.0), permute=True, seed=j + 1, as_particles=True) for j in range(16)]
    v = [mcm.lincomb([1.0 + 0.001 * j, 0.5], [base[j % 16], base[(j * 7 + 3) % 16]]) for j in range(M)]
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
    dC = mcm.particles._alloc(M * M * 8)
    res = []
    for mode in (0, 1):
        lib.mcm_gram_use_tensor(mode)
        L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None)); mcm.sync()
        ts = []
        for _ in range(3 if timing else 1):
            t = time.perf_counter(); L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None)); mcm.sync(); ts.append(time.perf_counter() - t)
        out = np.empty((M, M)); L.check(lib.mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(dC), out.nbytes))
        res.append((out, min(ts)))
    lib.mcm_gram_use_tensor(0)
    a, b = res[0][0], res[1][0]
    d = np.sqrt(np.diag(a))
    rel = np.max(np.abs(a - b) / np.outer(d, d))
    print(f"M={M} N={n} off={offset}: max |tc - dmma| / (sigma_i sigma_j) = {rel:.3e}  sym={np.array_equal(b, b.T)}  "
          f"dmma {res[0][1]*1e3:.2f} ms ({2*M*M*n/res[0][1]/1e12:.1f} TF)  tcgen05 {res[1][1]*1e3:.2f} ms ({2*M*M*n/res[1][1]/1e12:.1f} TF)", flush=True)
for (M, n) in [(256, 4096), (256, 5000), (300, 70001), (1024, 20000)]:
    run(M, n)
run(512, 10000, offset=1e6)
# heavy-tailed rows: t-distributed (ratio ~ 20-30, tensor path) and one extreme outlier (guard -> DMMA)
import numpy as _np
def heavy(M, n, outlier):
    rng = _np.random.default_rng(3)
    X = rng.standard_t(3, size=(M, n))
    if outlier: X[7, 123] = 1e6
    v = [mcm.Particles(_np.ascontiguousarray(X[j])) for j in range(M)]
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v]); dC = mcm.particles._alloc(M * M * 8)
    out = []
    for mode in (0, 1):
        lib.mcm_gram_use_tensor(mode)
        L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None))
        o = _np.empty((M, M)); L.check(lib.mcm_download(o.ctypes.data_as(C.c_void_p), C.c_void_p(dC), o.nbytes)); out.append((o, lib.mcm_gram_last_engine()))
    d = _np.sqrt(_np.diag(out[0][0]))
    ratio = _np.max(_np.abs(X - X.mean(axis=1, keepdims=True)), axis=1) / X.std(axis=1)
    print(f"heavy-tailed M={M} N={n} outlier={outlier}: max ratio {ratio.max():.1f}  engine {out[1][1]}  dev {_np.max(_np.abs(out[0][0]-out[1][0])/_np.outer(d,d)):.3e}", flush=True)
heavy(256, 20000, False); heavy(256, 20000, True)
run(4096, 100_000, timing=True)
run(1024, 1_000_000, timing=True)
