"""One transform_moments call at C5 size (launch list for ncu --metrics gpu__time_duration.sum)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
mcm.init(0)
d, n = 1024, 1_000_000
cols = mcm.systematic_batch(n, d, permute=True, seed=1)
rng = np.random.default_rng(0)
Aq = rng.normal(size=(d, d)); S = Aq @ Aq.T / d + np.eye(d); m = np.linspace(-1, 1, d)
Y = mcm.transform_moments(cols, m, S)
mcm.sync()
print("ok", float(mcm.pmean(Y[0])))
