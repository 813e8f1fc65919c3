"""A/B: Gram with the A operand read from shared memory (SS) vs staged in TMEM (TS)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcm_b200 as mcm
L = mcm._lib; lib = L.init(0)
for M, n in ((256, 8192), (512, 40000), (4096, 100000)):
    base = [mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=True, seed=j + 1, as_particles=True) for j in range(16)]
    v = [mcm.lincomb([1.0 + 0.001 * j, 0.5], [base[j % 16], base[(j * 7 + 3) % 16]]) for j in range(M)]
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
    dC = mcm.particles._alloc(M * M * 8)
    out = {}
    for mode in (1, 2):
        lib.mcm_gram_use_tensor(mode)
        ts = []
        for _ in range(3):
            mcm.sync(); t = time.perf_counter(); L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None)); mcm.sync(); ts.append(time.perf_counter() - t)
        h = np.empty((M, M)); L.check(lib.mcm_download(h.ctypes.data_as(C.c_void_p), C.c_void_p(dC), h.nbytes))
        out[mode] = (min(ts), h)
    lib.mcm_gram_use_tensor(1)
    d = np.sqrt(np.diag(out[1][1]))
    # mode 1 = library default (A operand staged in TMEM when g_a_in_tmem), mode 2 = A read from shared memory
    print(f"M={M} N={n}: mode1 (A in TMEM) {out[1][0]*1e3:.2f} ms ({2*M*M*n/out[1][0]/1e12:.1f} TF)  mode2 (A from smem) {out[2][0]*1e3:.2f} ms ({2*M*M*n/out[2][0]/1e12:.1f} TF)  "
          f"max |diff| / (sigma_i sigma_j) = {np.max(np.abs(out[2][1] - out[1][1]) / np.outer(d, d)):.3e}  bit-equal={np.array_equal(out[1][1], out[2][1])}")
