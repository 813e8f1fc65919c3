"""C5: d = 1024 uncertain variables x N = 1e6 particles, Float64, end to end on one B200:
   build (d keyed systematic columns)  ->  transform_moments to (m, S)  ->  bootstrap with one shared index
   vector  ->  pquantile at {0.025, 0.5, 0.975} of every variable.
Every stage is timed with wall clock around a synchronising call (best of k) and checked.
    python scripts/bench_c5.py [d] [n]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
GB = 1e9


def best(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    mcm.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); mcm.sync(); ts.append(time.perf_counter() - t)
    return min(ts), r


rng = np.random.default_rng(0)
Aq = rng.normal(size=(d, d))
S = Aq @ Aq.T / d + np.eye(d)
m = np.linspace(-1, 1, d)
total = {}

# 1. build
t, cols = best(lambda: mcm.systematic_batch(n, d, permute=True, seed=1))
total["build"] = t
print(f"C5 build: {d} keyed systematic columns, N={n}: {t*1e3:.2f} ms  {d*n/t/1e9:.1f} Gsamples/s  ({8*d*n/t/GB:.0f} GB/s written)")
ptrs = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in cols])

# 2. transform_moments, stage by stage and as one call
means = np.empty(d)
t_mu, _ = best(lambda: L.check(lib.mcm_colmeans(L.F64, ptrs, d, n, means.ctypes.data_as(L._dp))))
dC = mcm.particles._alloc(d * d * 8)
t_gram, _ = best(lambda: L.check(lib.mcm_gram(L.F64, ptrs, d, n, 1, None, C.c_void_p(dC), None)))
eng = lib.mcm_gram_last_engine()
cov = mcm.pcov(cols)
t0 = time.perf_counter()
xl = np.linalg.cholesky(cov)
A = np.tril(np.linalg.solve(xl.T, np.linalg.cholesky(S).T).T)
t_host = time.perf_counter() - t0
outs = [mcm.particles._empty(n, L.F64) for _ in range(d)]
optr = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in outs])
Ac = np.ascontiguousarray(A)
Af = np.ascontiguousarray(A + np.triu(np.full((d, d), 1e-300), 1))     # same values, no structural zeros
apply = lambda M_: L.check(lib.mcm_pgemm(L.F64, M_.ctypes.data_as(L._dp), d, d, ptrs, means.ctypes.data_as(L._dp),
                                        m.ctypes.data_as(L._dp), optr, n))
res_apply = {}
for eng, name in ((1, "tcgen05 int8 (7-digit Ozaki)"), (0, "FP64 DMMA")):
    lib.mcm_pgemm_use_tensor(eng)
    t_tri, _ = best(lambda: apply(Ac), reps=3)
    used = lib.mcm_pgemm_last_engine()
    t_full, _ = best(lambda: apply(Af), reps=2)
    res_apply[eng] = (t_tri, t_full)
    print(f"C5 apply Y = m + A (X - mu), {name} (engine used {used}): lower-triangular A {t_tri*1e3:.2f} ms "
          f"({d*d*n/t_tri/1e12:.1f} TFLOP/s executed) | dense A {t_full*1e3:.2f} ms ({2*d*d*n/t_full/1e12:.1f} TFLOP/s)")
lib.mcm_pgemm_use_tensor(1)
print(f"C5 transform_moments pieces: colmeans {t_mu*1e3:.2f} ms | Gram {t_gram*1e3:.2f} ms ({2*d*d*n/t_gram/1e12:.1f} TFLOP/s, engine {eng}) | "
      f"host 2x potrf + solve {t_host*1e3:.1f} ms")
t_tm, Y = best(lambda: mcm.transform_moments(cols, m, S), reps=2)
total["transform_moments"] = t_tm
print(f"C5 transform_moments, one call (colmeans + Gram + host factorisations + apply): {t_tm*1e3:.1f} ms")
pm_err = np.max(np.abs(np.array(mcm.pmean(Y)) - m))
pc_err = np.max(np.abs(mcm.pcov(Y) - S))
print(f"   check: max|pmean(Y) - m| = {pm_err:.2e}, max|pcov(Y) - S| = {pc_err:.2e}")
del cols, outs

# 3. bootstrap, one shared index vector
t_idx, idx = best(lambda: mcm.rand_indices(n, n, seed=3))
yp = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in Y])
res = [mcm.particles._empty(n, L.F64) for _ in range(d)]
rp = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in res])
t_g, _ = best(lambda: L.check(lib.mcm_gather_multi(L.F64, yp, rp, d, C.c_void_p(idx.ptr), n, n)))
total["bootstrap"] = t_idx + t_g
print(f"C5 bootstrap: indices {t_idx*1e3:.3f} ms | shared-index gather of {d} variables {t_g*1e3:.2f} ms  "
      f"{16*d*n/t_g/GB:.0f} GB/s algorithmic (8 B read + 8 B write per element)")
ih = idx.to_numpy()
for j in (0, d // 2, d - 1):
    assert np.array_equal(res[j].to_numpy(), Y[j].to_numpy()[ih]), j

# 4. quantiles of every variable
q = np.array([0.025, 0.5, 0.975])
t_q, Q = best(lambda: mcm.pquantile(res, q))
total["pquantile"] = t_q
print(f"C5 pquantile: {d} variables x 3 probabilities: {t_q*1e3:.2f} ms  {8*d*n/t_q/GB:.0f} GB/s per 8 B/key read-equivalent")
for j in (0, 1, d // 2, d - 1):
    ref = np.quantile(res[j].to_numpy(), q)
    assert np.allclose(Q[j], ref, rtol=1e-15, atol=0), (j, Q[j], ref)
t1, q1 = best(lambda: mcm.pquantile(res[0], q))
assert np.array_equal(q1, Q[0])
print(f"   (single-variable path: {t1*1e3:.3f} ms per variable -> {t1*d*1e3:.0f} ms for all)")
print("C5 total: " + " + ".join(f"{k} {v*1e3:.1f}" for k, v in total.items()) + f" = {sum(total.values())*1e3:.1f} ms")
