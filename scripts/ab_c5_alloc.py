import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
import ctypes as C
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
stream = torch.cuda.Stream()
L.check(lib.mcm_set_stream(C.c_void_p(stream.cuda_stream)))
d, n = 1024, 1_000_000
rng = np.random.default_rng(0)
Aq = rng.normal(size=(d, d)); S = Aq @ Aq.T / d + np.eye(d); m = np.linspace(-1, 1, d)
q = np.array([0.025, 0.5, 0.975])
def timed(name, fn, reps=2):
    res = None; ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize(); t = time.perf_counter(); res = fn(); mcm.sync(); ts.append((time.perf_counter() - t) * 1e3)
    print(name, " ".join("%.2f" % v for v in ts), flush=True)
    return res
with torch.cuda.stream(stream):
    cols = timed("build", lambda: mcm.systematic_batch(n, d, permute=True, seed=1))
    Y = timed("transform", lambda: mcm.transform_moments(cols, m, S))
    del cols
    R = timed("bootstrap", lambda: mcm.bootstrap(Y, n, seed=3), reps=4)
    Q = timed("pquantile", lambda: mcm.pquantile(R, q), reps=3)
    free, total = torch.cuda.mem_get_info(); print("free GB", free / 1e9, "total", total / 1e9)
