import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
M = int(sys.argv[1]); n = int(float(sys.argv[2])); reps = int(sys.argv[3])
base = [mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=True, seed=j + 1, as_particles=True) for j in range(16)]
v = [mcm.lincomb([1.0 + 0.001 * j, 0.5], [base[j % 16], base[(j * 7 + 3) % 16]]) for j in range(M)]
ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
dC = mcm.particles._alloc(M * M * 8)
lib.mcm_gram_use_tensor(1)
for _ in range(reps):
    t = time.perf_counter(); L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None)); mcm.sync()
    print(f"tc gram M={M} N={n}: {(time.perf_counter()-t)*1e3:.2f} ms")
