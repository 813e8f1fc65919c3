"""Run the dense tests and one transform_moments with the second-generation dense steps switched on."""
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import mcm_b200 as mcm
mcm.init(0)
lib = mcm._lib.lib()
lib.mcm_dense_use_v2(1)
ok = True
for M in (1, 5, 64, 65, 200, 517, 1024):
    rng = np.random.default_rng(M)
    Q = rng.normal(size=(M, M + 3)); A = Q @ Q.T / M + 0.5 * np.eye(M)
    Ld = mcm.cholesky_lower(A); Lh = np.linalg.cholesky(A)
    e1 = np.max(np.abs(Ld - Lh)) / (M * np.abs(Lh).max())
    B = np.tril(rng.normal(size=(M, M)))
    X = mcm.solve_right_lower(Lh, B)
    e2 = np.max(np.abs(X @ Lh - B)) / (M * max(1.0, np.abs(X).max()))
    good = e1 <= 1e-13 and e2 <= 1e-12 and np.all(np.triu(Ld, 1) == 0)
    ok &= bool(good)
    print(f"M={M}: potrf err {e1:.1e} trsm residual {e2:.1e} {'ok' if good else 'FAIL'}")
try:
    Bad = A.copy(); Bad[M // 2, M // 2] = -1.0
    mcm.cholesky_lower(Bad); ok = False; print("non-PD matrix was not rejected")
except mcm.MCMError:
    print("non-PD rejected ok")
d, n = 256, 20000
cols = mcm.systematic_batch(n, d, permute=True, seed=5)
rng = np.random.default_rng(3); Aq = rng.normal(size=(d, d)); S = Aq @ Aq.T / d + np.eye(d); m = np.linspace(-2, 2, d)
Y1 = mcm.transform_moments(cols, m, S)
Y1b = mcm.transform_moments(cols, m, S)
lib.mcm_dense_use_v2(0)
Y0 = mcm.transform_moments(cols, m, S)
lib.mcm_dense_use_v2(1)
err = max(np.max(np.abs(Y1[j].to_numpy() - Y0[j].to_numpy())) for j in (0, d // 2, d - 1))
cerr = np.max(np.abs(mcm.pcov(Y1) - S))
rep = np.array_equal(Y1[7].to_numpy(), Y1b[7].to_numpy())
print(f"transform_moments v2 vs v1: max diff {err:.1e}; max|pcov(Y) - S| {cerr:.1e}; repeatable {rep}")
ok &= err < 1e-10 and cerr < 1e-11 and rep
import time
d, n = 1024, 1_000_000
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
Aq = rng.normal(size=(d, d)); S = Aq @ Aq.T / d + np.eye(d); m = np.linspace(-1, 1, d)
for mode in (0, 1, 0, 1):
    lib.mcm_dense_use_v2(mode)
    mcm.transform_moments(cols, m, S); mcm.sync()
    t = time.perf_counter(); Y = mcm.transform_moments(cols, m, S); mcm.sync(); dt = time.perf_counter() - t
    print(f"C5 transform_moments dense v{mode + 1}: {dt*1e3:.1f} ms  max|pcov(Y)-S| {np.max(np.abs(mcm.pcov(Y) - S)):.1e}")
    del Y
print("V2CHECK", "PASS" if ok else "FAIL")
