"""Driver for profiling psort / psortperm at N = 1e8."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
x = mcm.rand_particles(n, seed=1)
for _ in range(2):
    mcm.sync(); t = time.perf_counter(); s = mcm.psort(x); mcm.sync(); print(f"psort N={n}: {(time.perf_counter()-t)*1e3:.2f} ms  {n/(time.perf_counter()-t)/1e9:.2f} Gkeys/s")
for _ in range(2):
    mcm.sync(); t = time.perf_counter(); p = mcm.psortperm(x); mcm.sync(); print(f"psortperm N={n}: {(time.perf_counter()-t)*1e3:.2f} ms")
x32 = x.astype(np.float32)
for _ in range(2):
    mcm.sync(); t = time.perf_counter(); s = mcm.psort(x32); mcm.sync(); print(f"psort f32 N={n}: {(time.perf_counter()-t)*1e3:.2f} ms")
