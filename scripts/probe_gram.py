"""Probe: pcov Gram kernel at a given size (for ncu / timing)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dt = L.F32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else L.F64
base = [mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=True, seed=j + 1, as_particles=True) for j in range(16)]
v = [mcm.lincomb([1.0 + 0.001 * j, 0.5], [base[j % 16], base[(j * 7 + 3) % 16]]) for j in range(M)]
if dt == L.F32:
    v = [p.astype(np.float32) for p in v]
ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
dC = mcm.particles._alloc(M * M * 8)
for r in range(reps):
    t = time.perf_counter(); L.check(lib.mcm_gram(dt, ptrs, M, n, 1, None, C.c_void_p(dC), None)); mcm.sync()
    e = time.perf_counter() - t
    print(f"gram M={M} N={n}: {e*1e3:.2f} ms  {2*M*M*n/e/1e12:.2f} TFLOP/s (GEMM convention)")
