import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import mcm_b200 as mcm
from mcm_b200 import particles as P
L = mcm._lib; lib = L.init(0)
d, n = 1024, 1_000_000
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
mcm.sync()
for mode in ("slab", "separate", "slab", "separate"):
    if mode == "separate":
        orig = P._empty_many
        P._empty_many = lambda M, N, dt: [P._empty(N, dt) for _ in range(M)]
    R = None
    ts = []
    for _ in range(5):
        mcm.sync(); t = time.perf_counter(); R = mcm.bootstrap(cols, n, seed=3); mcm.sync(); ts.append((time.perf_counter() - t) * 1e3)
    print(mode, " ".join("%.2f" % v for v in ts))
    # phases
    mcm.sync(); t = time.perf_counter(); outs = P._empty_many(d, n, cols[0].dtype); t1 = time.perf_counter(); mcm.sync(); t2 = time.perf_counter()
    print("   alloc %.2f ms (+sync %.2f)" % ((t1 - t) * 1e3, (t2 - t1) * 1e3))
    del outs, R
    if mode == "separate":
        P._empty_many = orig
