"""Time mcm_batched_linalg: N particles of n x n matrices, every operation; HBM bytes = 8 (inputs + outputs) per particle.
    python scripts/probe_linalg.py [N] [sizes, e.g. 12,16]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
mcm.init(0)
la = mcm.linalg
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000


def best(fn, reps=3):
    fn(); mcm.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); mcm.sync(); ts.append(time.perf_counter() - t)
    return min(ts)


SIZES = tuple(int(v) for v in sys.argv[2].split(',')) if len(sys.argv) > 2 else (2, 3, 4, 6, 8)
for n in SIZES:
    A = [[mcm.systematic_sample(N, mcm.Normal(1.0 if i == j else 0.2, 0.05), seed=10 * i + j, as_particles=True)
          for j in range(n)] for i in range(n)]
    b = [A[i][0] for i in range(n)]
    ops = [("solve (1 rhs)", lambda: la.solve(A, b), n * n + 2 * n), ("inv", lambda: la.inv(A), 2 * n * n),
           ("det", lambda: la.det(A), n * n + 1), ("exp", lambda: la.expm(A), 2 * n * n),
           ("eigvals", lambda: la.eigvals(A), n * n + 2 * n), ("eigvals_sym", lambda: la.eigvals_sym(A), n * n + n),
           ("svdvals", lambda: la.svdvals(A), n * n + n)]
    if n <= 4:
        ops.append(("lyap", lambda: la.lyap(A, A), 3 * n * n))
    line = []
    for name, fn, words in ops:
        t = best(fn)
        line.append(f"{name} {t*1e3:.2f} ms ({N/t/1e6:.0f} M/s, {8*words*N/t/1e9:.0f} GB/s)")
    print(f"n={n} N={N}: " + " | ".join(line), flush=True)
    del A, b
# CPU yardstick: the reference's execution model is a host loop of LAPACK calls, one per particle
import scipy.linalg as sl
M = np.random.default_rng(0).normal(size=(20000, 3, 3))
t = time.perf_counter(); [np.linalg.eigvals(m) for m in M]; te = time.perf_counter() - t
t = time.perf_counter(); [sl.expm(m) for m in M]; tx = time.perf_counter() - t
t = time.perf_counter(); np.linalg.inv(M); ti = time.perf_counter() - t
print(f"host loop, 3x3, 1 core: eigvals {len(M)/te/1e6:.3f} M/s, scipy expm {len(M)/tx/1e6:.3f} M/s, numpy batched inv {len(M)/ti/1e6:.2f} M/s")
