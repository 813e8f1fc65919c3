"""Time the batched systematic build (C5 size) in scatter and gather form; check they agree."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
d, n = 1024, 1_000_000
res = {}
for mode, name in ((1, "scatter"), (0, "gather")):
    lib.mcm_systematic_use_scatter(mode)
    ts = []
    for _ in range(3):
        mcm.sync(); t = time.perf_counter(); cols = mcm.systematic_batch(n, d, permute=True, seed=1); mcm.sync(); ts.append(time.perf_counter() - t)
    res[mode] = [cols[j].to_numpy() for j in (0, 517, 1023)]
    print(f"build {d} x {n} {name}: {min(ts)*1e3:.2f} ms")
    del cols
lib.mcm_systematic_use_scatter(1)
print("identical:", all(np.array_equal(a, b) for a, b in zip(res[0], res[1])))
