"""mcm_sort / mcm_sortperm at N = 1e8: CUDA-event time of the C call alone, for several data sets."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import mcm_b200 as mcm
from mcm_b200 import particles as P
L = mcm._lib; lib = L.init(0)
stream = torch.cuda.Stream()
L.check(lib.mcm_set_stream(C.c_void_p(stream.cuda_stream)))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
data = {"N(0,1) iid": mcm.rand_particles(n, seed=1), "N(pi,0.1) systematic, permuted": mcm.systematic_sample(n, mcm.Normal(np.pi, 0.1), permute=True, seed=3, as_particles=True)}
out = P._empty(n, L.F64)
perm = P.DeviceIndex(n)
for name, x in data.items():
    for what in ("sort", "sortperm") + (("sort-unfused", "sortperm-unfused") if hasattr(lib, "mcm_sort_use_fused") else ()):
        ts = []
        with torch.cuda.stream(stream):
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); t = time.perf_counter()
                e0.record(stream)
                lib.mcm_sort_use_fused(0 if what.endswith("unfused") else 1)
                if what.startswith("sort") and not what.startswith("sortperm"):
                    L.check(lib.mcm_sort(L.F64, C.c_void_p(x.ptr), C.c_void_p(out.ptr), n))
                else:
                    L.check(lib.mcm_sortperm(L.F64, C.c_void_p(x.ptr), C.c_void_p(perm.ptr), C.c_void_p(out.ptr), n))
                e1.record(stream); torch.cuda.synchronize()
                ts.append((e0.elapsed_time(e1), (time.perf_counter() - t) * 1e3))
        print("%-32s %-8s events %.2f ms  wall %.2f ms (best of 4)" % (name, what, min(t[0] for t in ts), min(t[1] for t in ts)), flush=True)
    s = out.to_numpy()[:: max(1, n // 1000)]
    assert np.all(np.diff(s) >= 0)
