"""Where the wall time of transform_moments at C5 size goes: Python wrapper vs the C call vs the kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import mcm_b200 as mcm
from mcm_b200 import particles as P
L = mcm._lib; lib = L.init(0)
d, n = 1024, 1_000_000
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
rng = np.random.default_rng(0)
Aq = rng.normal(size=(d, d)); S = np.ascontiguousarray(Aq @ Aq.T / d + np.eye(d)); m = np.ascontiguousarray(np.linspace(-1, 1, d))
for _ in range(2):
    Y = mcm.transform_moments(cols, m, S); mcm.sync()
for rep in range(3):
    mcm.sync(); t0 = time.perf_counter()
    outs = P._empty_many(d, n, cols[0].dtype)
    pin = (C.c_void_p * d)(*[C.c_void_p(x.ptr) for x in cols])
    pout = (C.c_void_p * d)(*[C.c_void_p(y.ptr) for y in outs])
    t1 = time.perf_counter()
    L.check(lib.mcm_transform_moments(cols[0].dtype, pin, d, n, m.ctypes.data_as(L._dp), S.ctypes.data_as(L._dp), 0, pout))
    t2 = time.perf_counter()
    mcm.sync(); t3 = time.perf_counter()
    print("wrapper prep %.2f ms | C call returns after %.2f ms | sync +%.2f ms | total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
    del outs
