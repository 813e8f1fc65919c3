"""A few launches of one generated chain kernel, for ncu:
    ncu --set full --import-source on --clock-control none -k regex:mcm_jit_kernel -s 3 -c 1 -o gpurun_out/x python scripts/prof_chain.py [chain10|rk4] [nostats]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcm_b200 as mcm
from mcm_b200 import chains, tape as tp

L = mcm._lib
mcm.init(0)
lib = L.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "chain10"
stats = "nostats" not in sys.argv
for a in sys.argv[2:]:
    if a.startswith("opt"):          # opt<which>=<value>
        w, v = a[3:].split("=")
        L.check(lib.mcm_chain_jit_option(int(w), int(v)))
if which == "chain10":
    N = 100_000_000
    ins = [mcm.systematic_sample(N, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)]
    outs = [mcm.particles._empty(N, L.F64)]
    t = tp.trace(chains.chain10, 1)
    reps = 0
else:
    N = 10_000_000
    ins = [mcm.systematic_sample(N, d, permute=True, seed=11 + k, as_particles=True)
           for k, d in enumerate((mcm.Normal(0.0, 1e-12), mcm.Normal(np.pi / 3, 0.02), mcm.Normal(9.79, 0.02), mcm.Normal(1.0, 0.01), mcm.Normal(0.1, 0.01)))]
    outs = [mcm.particles._empty(N, L.F64) for _ in range(2)]
    t = chains.pendulum_rk4_tape(0.01)
    reps = 100
insp = (C.c_void_p * len(ins))(*[p.ptr for p in ins])
outp = (C.c_void_p * len(outs))(*[p.ptr for p in outs])
consts = np.asarray(t.consts if t.consts else [0.0])
tc = t.as_ctypes(); lb, le = t.bounds
for _ in range(6):
    L.check(lib.mcm_chain(L.F64, tc, len(t.ins), lb, le, reps, consts.ctypes.data_as(L._dp), len(t.consts), insp, len(ins),
                          outp, len(outs), N, (t.out_slots[0] if which == "chain10" else 0) if stats else -1, None))
L.check(lib.mcm_sync())
print("engine", lib.mcm_chain_last_engine())
