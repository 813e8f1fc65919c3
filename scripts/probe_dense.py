"""Time the M x M dense device routines (potrf, trsm) against host LAPACK."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
for M in (256, 1024, 2048):
    rng = np.random.default_rng(M)
    Q = rng.normal(size=(M, M)); A = Q @ Q.T / M + np.eye(M)
    dA = mcm.particles._alloc(A.nbytes); dB = mcm.particles._alloc(A.nbytes)
    info = C.c_int(0)
    ts = []
    for _ in range(3):
        L.check(lib.mcm_upload(C.c_void_p(dA), A.ctypes.data_as(C.c_void_p), A.nbytes)); mcm.sync()
        t = time.perf_counter(); L.check(lib.mcm_potrf(C.c_void_p(dA), M, C.byref(info))); mcm.sync(); ts.append(time.perf_counter() - t)
    Bm = np.tril(rng.normal(size=(M, M)))
    tt = []
    for _ in range(3):
        L.check(lib.mcm_upload(C.c_void_p(dB), Bm.ctypes.data_as(C.c_void_p), Bm.nbytes)); mcm.sync()
        t = time.perf_counter(); L.check(lib.mcm_trsm_right_lower(C.c_void_p(dA), C.c_void_p(dB), M, M)); mcm.sync(); tt.append(time.perf_counter() - t)
    t = time.perf_counter(); Lh = np.linalg.cholesky(A); th = time.perf_counter() - t
    t = time.perf_counter(); Xh = np.linalg.solve(Lh.T, Bm.T).T; ts2 = time.perf_counter() - t
    print(f"M={M}: device potrf {min(ts)*1e3:.2f} ms, trsm {min(tt)*1e3:.2f} ms | host LAPACK potrf {th*1e3:.1f} ms, solve {ts2*1e3:.1f} ms")
