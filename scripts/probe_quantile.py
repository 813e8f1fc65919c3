"""Probe: quantile select vs sort paths (exactness + timing)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
rng = np.random.default_rng(0)
q = np.array([0.0, 0.025, 0.25, 0.5, 0.975, 1.0, 1 / 3, 0.9999])
def both(p):
    r = []
    for mode in (0, 1):
        lib.mcm_quantile_use_select(mode)
        r.append(mcm.pquantile(p, q))
    return r
for name, x in (("normal 1e6", rng.normal(size=1_000_003)), ("constant", np.full(100_001, 2.5)), ("two values", np.repeat([1.0, -3.0], 60_000)),
                ("tiny n=1", np.array([4.0])), ("n=2", np.array([4.0, -1.0])), ("n=5000", rng.normal(size=5000)), ("ints", rng.integers(0, 50, 300_000).astype(np.float64)),
                ("wide dyn range", np.exp(rng.uniform(-300, 300, 500_000)) * rng.choice([-1, 1], 500_000)), ("with inf", np.concatenate([rng.normal(size=1000), [np.inf, -np.inf]])),
                ("f32", rng.normal(size=700_001).astype(np.float32)), ("signed zeros", np.concatenate([np.zeros(1000), -np.zeros(1000)]))):
    p = mcm.Particles(x)
    a, b = both(p)
    ref = np.quantile(x.astype(np.float64), q)
    print(f"{name:16s} select==sort: {np.array_equal(a, b)}  max|sel-numpy|={np.max(np.abs(np.nan_to_num(b - ref, nan=0, posinf=0, neginf=0))):.2e}")
try:
    lib.mcm_quantile_use_select(1); mcm.pquantile(mcm.Particles(np.array([1.0, np.nan, 2.0])), 0.5); print("NaN: no error (BAD)")
except mcm.ArgumentError as e:
    print("NaN -> ArgumentError ok")
for n in (1_000_000, 10_000_000, 100_000_000):
    x = mcm.systematic_sample(n, mcm.Normal(), permute=True, seed=5, as_particles=True)
    q3 = np.array([0.025, 0.5, 0.975]); out = np.zeros(3)
    for mode in (0, 1):
        lib.mcm_quantile_use_select(mode)
        ts = []
        for _ in range(4):
            t = time.perf_counter(); L.check(lib.mcm_quantile(L.F64, C.c_void_p(x.ptr), n, q3.ctypes.data_as(L._dp), 3, out.ctypes.data_as(L._dp))); ts.append(time.perf_counter() - t)
        print(f"N={n} select={mode}: {min(ts)*1e3:.3f} ms  {8*n/min(ts)/1e9:.0f} GB/s of keys  {out}")
