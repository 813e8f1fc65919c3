"""Driver for profiling mcm_pgemm at C5 size (d x d matrix, d variables x n particles)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
tri = (sys.argv[3] if len(sys.argv) > 3 else "tri") == "tri"
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
rng = np.random.default_rng(0)
A = rng.normal(size=(d, d)) / np.sqrt(d)
if tri: A = np.tril(A)
A = np.ascontiguousarray(A)
ptrs = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in cols])
outs = [mcm.particles._empty(n, L.F64) for _ in range(d)]
optr = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in outs])
mu = np.zeros(d); b = np.zeros(d)
for eng in (1, 0):
    lib.mcm_pgemm_use_tensor(eng)
    for _ in range(2):
        mcm.sync(); t = time.perf_counter()
        L.check(lib.mcm_pgemm(L.F64, A.ctypes.data_as(L._dp), d, d, ptrs, mu.ctypes.data_as(L._dp), b.ctypes.data_as(L._dp), optr, n))
        mcm.sync(); print("engine", lib.mcm_pgemm_last_engine(), "ms", (time.perf_counter() - t) * 1e3)
lib.mcm_pgemm_use_tensor(1)
