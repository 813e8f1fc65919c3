"""Driver for profiling the batched quantile / build / gather kernels at C5 size (used under ncu)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
q = np.array([0.025, 0.5, 0.975])
for _ in range(2):
    mcm.sync(); t = time.perf_counter(); Q = mcm.pquantile(cols, q); mcm.sync(); print("pquantile ms", (time.perf_counter() - t) * 1e3)
print(Q[0], "fallbacks", lib.mcm_quantile_batch_fallbacks())
big = mcm.systematic_batch(100_000_000, 1, permute=True, seed=2)
for _ in range(2):
    mcm.sync(); t = time.perf_counter(); Q1 = mcm.pquantile(big, q); mcm.sync(); print("pquantile batch path M=1 N=1e8 ms", (time.perf_counter() - t) * 1e3)
for mode, name in ((1, "default (brackets)"), (2, "MSD radix select"), (0, "full sort")):
    lib.mcm_quantile_use_select(mode)
    for _ in range(2):
        mcm.sync(); t = time.perf_counter(); Q2 = mcm.pquantile(big[0], q); mcm.sync(); print(f"pquantile single N=1e8, {name}: ms", (time.perf_counter() - t) * 1e3)
    assert np.array_equal(Q1[0], Q2)
lib.mcm_quantile_use_select(1)
