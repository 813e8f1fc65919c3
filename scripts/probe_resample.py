"""Time systematic resampling at N = 1e8 (expansion vs per-output search)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
z = mcm.rand_particles(n, seed=1)
for name, w in (("exp(z)", mcm.exp(z)), ("exp(8 z) (peaked)", mcm.exp(8.0 * z)), ("flat", mcm.fill(1.0, n))):
    idx = mcm.DeviceIndex(n)
    for mode, mname in ((0, "expansion"), (1, "search")):
        lib.mcm_resample_use_search(mode)
        ts = []
        for _ in range(3):
            mcm.sync(); t = time.perf_counter()
            L.check(lib.mcm_resample_systematic(C.c_void_p(w.ptr), n, 0.37, C.c_void_p(idx.ptr))); mcm.sync()
            ts.append(time.perf_counter() - t)
        print(f"resample_systematic N={n} weights {name}: {mname} {min(ts)*1e3:.2f} ms")
lib.mcm_resample_use_search(0)
