"""Generate tests/golden/norminv_truth.npz (run here, in the build container):
    python tests/golden/make_norminv_golden.py

Truth for quantile(Normal(0,1), p): Newton on ln Phi(z) = ln p in mpmath at 60 digits, stored as a
double-double (y, y_lo) so that errors of a Float64 implementation are resolved to ~0.001 ulp.  Points:
the systematic grids of N = 100 / 2000 (src/sampling.jl:8-10), uniform random p, both tails down to the
subnormals, p next to 1/2, p next to 1, and p around every branch / piece boundary of csrc/mcm_norminv.cuh."""
import os

import mpmath as mp
import numpy as np
from scipy.special import ndtri_exp

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 60
rng = np.random.default_rng(20261020)


def ppf(pf):
    p = mp.mpf(float(pf))
    flip = p > mp.mpf(1) / 2
    if flip:
        p = 1 - p
    lp = mp.log(p)
    z = mp.mpf(float(ndtri_exp(float(lp))))
    for _ in range(4):
        c = mp.ncdf(z)
        z = z - (mp.log(c) - lp) * c / mp.npdf(z)
    return -z if flip else z


parts = [
    (2 * np.arange(2000) + 1) / 4000.0,
    (2 * np.arange(100) + 1) / 200.0,
    rng.random(12000),
    10.0 ** (-rng.uniform(0, 323, 4000)),
    1 - 10.0 ** (-rng.uniform(0, 16, 2000)),
    0.5 + rng.choice([-1, 1], 2000) * 10.0 ** (-rng.uniform(0.3, 17, 2000)),
    0.125 + rng.normal(0, 1e-3, 500), 0.875 + rng.normal(0, 1e-3, 500),
    np.array([0.125, 0.875, np.nextafter(0.125, 0), np.nextafter(0.125, 1), np.nextafter(0.875, 0), np.nextafter(0.875, 1),
              5e-324, 2.2250738585072014e-308, np.nextafter(1.0, 0), np.nextafter(0.5, 0), np.nextafter(0.5, 1), 0.25, 0.75]),
]
for rb in (3, 4, 6, 8, 12, 16, 24, 32):          # piece boundaries of r = sqrt(-2 ln p)
    r = rb + rng.normal(0, 1e-3 * rb, 250)
    parts.append(np.exp(-r * r / 2))
p = np.concatenate(parts)
p = p[(p > 0) & (p < 1)]
hi = np.empty_like(p); lo = np.empty_like(p)
for i, v in enumerate(p):
    z = ppf(v)
    hi[i] = float(z)
    lo[i] = float(z - mp.mpf(hi[i]))
np.savez_compressed(os.path.join(HERE, "norminv_truth.npz"), p=p, y=hi, y_lo=lo)
print("norminv truth:", p.size, "points")
