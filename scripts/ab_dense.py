"""Time the device Cholesky (mcm_potrf) for the three generations of the diagonal-block kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import mcm_b200 as mcm
from mcm_b200 import particles as P
L = mcm._lib; lib = L.init(0)
rng = np.random.default_rng(0)
for M in (1024, 2048):
    Q = rng.normal(size=(M, M + 8)); A = np.ascontiguousarray(Q @ Q.T + np.eye(M))
    d = P._alloc(A.nbytes)
    info = C.c_int(0)
    for mode in (2, 1, 0, 2, 1):
        lib.mcm_dense_use_v2(mode)
        ts = []
        for _ in range(4):
            L.check(lib.mcm_upload(C.c_void_p(d), A.ctypes.data_as(C.c_void_p), A.nbytes)); mcm.sync()
            t = time.perf_counter(); L.check(lib.mcm_potrf(C.c_void_p(d), M, C.byref(info))); mcm.sync(); ts.append((time.perf_counter() - t) * 1e3)
        print("M=%d mode %d: potrf %.3f ms (best of 4)" % (M, mode, min(ts)), flush=True)
    lib.mcm_dense_use_v2(2)
