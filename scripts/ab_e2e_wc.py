"""e2e step (mcm_chain_host, 800 MB of pinned host input per step) with default pinned vs write-combined pinned input."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import mcm_b200 as mcm
from mcm_b200 import chains, tape as tp
L = mcm._lib; lib = L.init(0)
N = 100_000_000
a = mcm.systematic_sample(N, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)
t10 = tp.trace(chains.chain10, 1)
consts = np.asarray(t10.consts if t10.consts else [0.0]); lb, le = t10.bounds; tc = t10.as_ctypes()
st = np.zeros(6)
for name, alloc in (("pinned", lib.mcm_host_alloc), ("write-combined", lib.mcm_host_alloc_wc), ("pinned", lib.mcm_host_alloc)):
    hp = C.c_void_p()
    L.check(alloc(C.byref(hp), N * 8))
    L.check(lib.mcm_download(hp, C.c_void_p(a.ptr), N * 8)); mcm.sync()
    hin = (C.c_void_p * 1)(hp.value)
    ts = []
    for _ in range(6):
        t = time.perf_counter()
        L.check(lib.mcm_chain_host(L.F64, tc, len(t10.ins), lb, le, 0, consts.ctypes.data_as(L._dp), len(t10.consts), hin, 1, None, 1, N,
                                   t10.out_slots[0], st.ctypes.data_as(L._dp)))
        ts.append(time.perf_counter() - t)
    print("%-15s best %.3f ms/step  %.1f GB/s H2D   mean %.15g" % (name, min(ts) * 1e3, 8 * N / min(ts) / 1e9, st[1]), flush=True)
    L.check(lib.mcm_host_free(hp))
