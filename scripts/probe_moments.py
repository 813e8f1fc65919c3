"""Probe: moments kernel A/B (bulk-async vs LDG) + correctness at 1e8."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
a = mcm.systematic_sample(N, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)
m = np.zeros(6)
for mode in (0, 1, 0, 1):
    lib.mcm_moments_use_bulk(mode)
    ts = []
    for _ in range(12):
        t = time.perf_counter(); L.check(lib.mcm_moments(L.F64, C.c_void_p(a.ptr), N, m.ctypes.data_as(L._dp))); ts.append(time.perf_counter() - t)
    best = min(ts)
    print(f"bulk={mode} N={N}: {best*1e6:.1f} us  {8*N/best/1e9:.0f} GB/s  mean={m[1]!r} std={np.sqrt(m[2]/(N-1))!r} min={m[3]} max={m[4]}")
for n in (2048 * 3 + 5, 1 << 20, (1 << 20) + 1, 10_000_001):
    x = np.random.default_rng(n).normal(3, 2, n)
    p = mcm.Particles(x)
    r = []
    for mode in (0, 1):
        lib.mcm_moments_use_bulk(mode)
        L.check(lib.mcm_moments(L.F64, C.c_void_p(p.ptr), n, m.ctypes.data_as(L._dp))); r.append(m.copy())
    ok = abs(r[0][1] - r[1][1]) < 1e-13 and abs(r[0][2] - r[1][2]) < 1e-12 * r[0][2] and r[0][3] == r[1][3] and r[0][4] == r[1][4] and r[0][0] == r[1][0] == n
    print("n", n, "A/B agree:", ok, abs(r[1][1] - x.mean()), abs(r[1][2] / (n - 1) - x.var(ddof=1)))
xf = np.random.default_rng(1).normal(3, 2, 5_000_003).astype(np.float32)
pf = mcm.Particles(xf)
lib.mcm_moments_use_bulk(1)
L.check(lib.mcm_moments(L.F32, C.c_void_p(pf.ptr), xf.size, m.ctypes.data_as(L._dp)))
print("f32 bulk", m[0] == xf.size, abs(m[1] - xf.astype(np.float64).mean()), m[3] == xf.min(), m[4] == xf.max())
