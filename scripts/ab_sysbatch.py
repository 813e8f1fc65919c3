import os, sys, time
sys.path.insert(0, "/root/repo")
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
for mode in (1, 2, 1, 2):
    lib.mcm_systematic_use_scatter(mode)
    for _ in range(2):
        mcm.sync(); t = time.perf_counter(); cols = mcm.systematic_batch(1_000_000, 1024, permute=True, seed=1); mcm.sync()
        print("mode", mode, "build %.2f ms" % ((time.perf_counter() - t) * 1e3))
    import numpy as np
    print("  check", float(mcm.pmean(cols[5])), float(mcm.pstd(cols[1000])))
