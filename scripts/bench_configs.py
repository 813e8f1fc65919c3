"""Measure the non-headline configurations of BASELINE.json (C3, C4, C5 pieces) and the supporting kernels
on one B200.  CUDA-event-free wall clock around synchronising calls (best of k); prints one line each.
    python scripts/bench_configs.py [c3] [c4] [c5] [sort] [misc]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
from mcm_b200 import chains, tape
L = mcm._lib; lib = L.init(0)
which = set(sys.argv[1:]) or {"c3", "c4", "c5", "sort", "misc"}


def best(fn, reps=5, warm=1):
    for _ in range(warm): fn()
    mcm.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); mcm.sync(); ts.append(time.perf_counter() - t)
    return min(ts)


if "c3" in which:
    n, steps = 10_000_000, 1000
    mk = lambda mu, s, seed: mcm.systematic_sample(n, mcm.Normal(mu, s), permute=True, seed=seed, as_particles=True)
    th, om = mcm.fill(0.0, n), mk(np.pi / 3, 0.02, 1)
    g, Lp, c = mk(9.79, 0.02, 2), mk(1.0, 0.01, 3), mk(0.1, 0.01, 4)
    t = chains.pendulum_rk4_tape(0.01)
    o1, o2 = mcm.particles._empty(n, L.F64), mcm.particles._empty(n, L.F64)
    dt = best(lambda: tape.run_tape(t, [p.ptr for p in (th, om, g, Lp, c)], n, L.F64, [o1.ptr, o2.ptr], reps=steps), reps=2)
    print(f"C3 pendulum RK4 fused in registers: N={n} steps={steps}: {dt*1e3:.1f} ms  {n*steps/dt/1e9:.2f} Gparticle-steps/s  "
          f"{44*n*steps/dt/1e12:.2f} Tparticle-ops/s  pmean(theta)={mcm.pmean(o1):.12g}")
    dt1 = best(lambda: tape.run_tape(t, [p.ptr for p in (th, om, g, Lp, c)], n, L.F64, [o1.ptr, o2.ptr], reps=1), reps=5)
    print(f"C3 one step per launch (trajectory saved): {dt1*1e3:.3f} ms/step  {56*n/dt1/1e9:.0f} GB/s algorithmic (56 B/particle/step)")

if "c4" in which:
    M, n = 4096, 100_000
    base = [mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=True, seed=j + 1, as_particles=True) for j in range(64)]
    v = [mcm.lincomb([1.0 + 0.001 * j, 0.5], [base[j % 64], base[(j * 7 + 3) % 64]]) for j in range(M)]
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
    dC = mcm.particles._alloc(M * M * 8)
    dt = best(lambda: L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None)), reps=5)
    print(f"C4 pcov M={M} N={n} F64 (DMMA): {dt*1e3:.2f} ms  {2*M*M*n/dt/1e12:.2f} TFLOP/s (GEMM convention; symmetric half computed)")
    vf = [p.astype(np.float32) for p in v]
    ptrf = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in vf])
    dt = best(lambda: L.check(lib.mcm_gram(L.F32, ptrf, M, n, 1, None, C.c_void_p(dC), None)), reps=5)
    print(f"C4 pcov M={M} N={n} F32 in, F64 accumulate (DMMA): {dt*1e3:.2f} ms  {2*M*M*n/dt/1e12:.2f} TFLOP/s")

if "c5" in which:
    d, n = 1024, 1_000_000
    t0 = time.perf_counter()
    cols = [mcm.systematic_sample(n, mcm.Normal(), permute=True, seed=1000003 * j, as_particles=True) for j in range(d)]
    mcm.sync(); t1 = time.perf_counter()
    print(f"C5 {d} systematic columns (randperm + Phi^-1 grid), N={n}: {(t1-t0)*1e3:.0f} ms  ({(t1-t0)/d*1e3:.2f} ms/column)")
    ptrs = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in cols])
    dC = mcm.particles._alloc(d * d * 8)
    dt = best(lambda: L.check(lib.mcm_gram(L.F64, ptrs, d, n, 1, None, C.c_void_p(dC), None)), reps=3)
    print(f"C5 Gram d={d} N={n}: {dt*1e3:.1f} ms  {2*d*d*n/dt/1e12:.2f} TFLOP/s")
    outs = [mcm.particles._empty(n, L.F64) for _ in range(d)]
    optr = (C.c_void_p * d)(*[C.c_void_p(p.ptr) for p in outs])
    idx = mcm.rand_indices(n, n, seed=3)
    dt = best(lambda: L.check(lib.mcm_gather_multi(L.F64, ptrs, optr, d, C.c_void_p(idx.ptr), n, n)), reps=3)
    print(f"C5 bootstrap gather {d} vars x {n}: {dt*1e3:.1f} ms  {16*d*n/dt/1e9:.0f} GB/s algorithmic")
    q = np.array([0.025, 0.5, 0.975]); out = np.zeros(3)
    t0 = time.perf_counter()
    for j in range(64):
        L.check(lib.mcm_quantile(L.F64, C.c_void_p(outs[j].ptr), n, q.ctypes.data_as(L._dp), 3, out.ctypes.data_as(L._dp)))
    dtq = (time.perf_counter() - t0) / 64
    print(f"C5 pquantile (3 probs) per variable N={n}: {dtq*1e3:.2f} ms  -> {dtq*d*1e3:.0f} ms for {d} variables  ({8*n/dtq/1e9:.0f} GB/s of keys)")

if "sort" in which:
    for n in (1_000_000, 10_000_000, 100_000_000):
        x = mcm.systematic_sample(n, mcm.Normal(), permute=True, seed=5, as_particles=True)
        o = mcm.particles._empty(n, L.F64)
        dt = best(lambda: L.check(lib.mcm_sort(L.F64, C.c_void_p(x.ptr), C.c_void_p(o.ptr), n)), reps=3)
        print(f"sort F64 N={n}: {dt*1e3:.2f} ms  {n/dt/1e9:.2f} Gkeys/s")
        pm = mcm.DeviceIndex(n)
        dt = best(lambda: L.check(lib.mcm_sortperm(L.F64, C.c_void_p(x.ptr), C.c_void_p(pm.ptr), None, n)), reps=3)
        print(f"sortperm F64 N={n}: {dt*1e3:.2f} ms  {n/dt/1e9:.2f} Gkeys/s")
        dt = best(lambda: L.check(lib.mcm_randperm(C.c_void_p(pm.ptr), n, 1)), reps=3)
        print(f"randperm N={n}: {dt*1e3:.2f} ms")
        q = np.array([0.025, 0.5, 0.975]); out = np.zeros(3)
        dt = best(lambda: L.check(lib.mcm_quantile(L.F64, C.c_void_p(x.ptr), n, q.ctypes.data_as(L._dp), 3, out.ctypes.data_as(L._dp))), reps=3)
        print(f"quantile F64 N={n}: {dt*1e3:.2f} ms  {8*n/dt/1e9:.0f} GB/s of keys")

if "misc" in which:
    n = 100_000_000
    o = mcm.particles._empty(n, L.F64)
    dt = best(lambda: L.check(lib.mcm_systematic_normal(L.F64, C.c_void_p(o.ptr), n, np.pi, 0.1, None, 0, n)))
    print(f"systematic_normal (Phi^-1 grid) N={n}: {dt*1e3:.3f} ms  {n/dt/1e9:.1f} Gsamples/s  {8*n/dt/1e9:.0f} GB/s written")
    idx = mcm.rand_indices(n, n, seed=3)
    dt = best(lambda: L.check(lib.mcm_rand_indices(C.c_void_p(idx.ptr), n, n, 3, 0)))
    print(f"rand_indices N={n}: {dt*1e3:.3f} ms  {8*n/dt/1e9:.0f} GB/s")
    g = mcm.particles._empty(n, L.F64)
    dt = best(lambda: L.check(lib.mcm_gather(L.F64, C.c_void_p(o.ptr), C.c_void_p(idx.ptr), C.c_void_p(g.ptr), n, n)))
    print(f"gather (uniform random) N={n}: {dt*1e3:.3f} ms  {24*n/dt/1e9:.0f} GB/s algorithmic (8 idx + 8 read + 8 write)")
    w = mcm.exp(0.5 * (o - np.pi) * 10.0)
    ri = mcm.DeviceIndex(n)
    dt = best(lambda: L.check(lib.mcm_resample_systematic(C.c_void_p(w.ptr), n, 0.5, C.c_void_p(ri.ptr))), reps=3)
    print(f"resample_systematic N={n}: {dt*1e3:.3f} ms")
    cnt = C.c_uint64()
    dt = best(lambda: L.check(lib.mcm_count(L.ENUMS['MCM_CMP_LT'], L.F64, C.c_void_p(o.ptr), n, 3.2, C.byref(cnt))))
    print(f"count (@prob) N={n}: {dt*1e3:.3f} ms  {8*n/dt/1e9:.0f} GB/s")
    xf = o.astype(np.float32); of = mcm.particles._empty(n, L.F32)
    for nm in ("SIN", "EXP", "SQRT"):
        dt = best(lambda: L.check(lib.mcm_unary(L.ENUMS['MCM_U_' + nm], L.F32, C.c_void_p(xf.ptr), C.c_void_p(of.ptr), n)))
        print(f"F32 {nm} N={n}: {dt*1e3:.3f} ms  {8*n/dt/1e9:.0f} GB/s")
    t10 = tape.trace(chains.chain10, 1)
    dt = best(lambda: tape.run_tape(t10, [xf.ptr], n, L.F32, [of.ptr], stats_output=0))
    print(f"F32 chain10+stats N={n}: {dt*1e3:.3f} ms  {10*n/dt/1e9:.0f} Gparticle-ops/s")
