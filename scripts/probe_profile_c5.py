"""One call of each C5 stage (for an ncu capture restricted by -k): build, transform application, quantiles, bootstrap."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
d, n = 1024, 1_000_000
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
rng = np.random.default_rng(0)
A = np.tril(rng.normal(size=(d, d))) / np.sqrt(d) + np.eye(d)
Y = mcm.pgemm(A, cols, shift=np.zeros(d), bias=np.linspace(-1, 1, d))
Q = mcm.pquantile(Y, np.array([0.025, 0.5, 0.975]))
srt = mcm.psort(Y[0])
B = mcm.bootstrap(Y, seed=3)
mcm.sync()
print("engine", lib.mcm_pgemm_last_engine(), "q0", Q[0], "boot", float(mcm.pmean(B[0])))
