"""A/B of the CTA size of the runtime-specialised chain kernels (mcm_chain_jit_option 5); derived from ab_jit.py.
(original docstring:) A/B of the generation tunables of the runtime-specialised chain kernels on a B200 (kernel-only CUDA-event times):
    python scripts/ab_jit.py > gpurun_out/ab_jit.txt"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mcm_b200 as mcm
from mcm_b200 import chains, tape as tp

L = mcm._lib
mcm.init(0)
lib = L.lib()
stream = torch.cuda.Stream()
L.check(lib.mcm_set_stream(C.c_void_p(stream.cuda_stream)))


def time_chain(t, ins, outs, n, reps_loop, stats_slot, k=20):
    insp = (C.c_void_p * len(ins))(*[p.ptr for p in ins])
    outp = (C.c_void_p * len(outs))(*[p.ptr for p in outs])
    consts = np.asarray(t.consts if t.consts else [0.0])
    tc = t.as_ctypes(); lb, le = t.bounds

    def go():
        L.check(lib.mcm_chain(L.F64, tc, len(t.ins), lb, le, reps_loop, consts.ctypes.data_as(L._dp), len(t.consts), insp, len(ins),
                              outp, len(outs), n, stats_slot, None))
    with torch.cuda.stream(stream):
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            go()
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, lib.mcm_chain_last_engine()


N = 100_000_000
a = mcm.systematic_sample(N, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)
y = mcm.particles._empty(N, L.F64)
t10 = tp.trace(chains.chain10, 1)
n3 = 10_000_000
ins3 = [mcm.systematic_sample(n3, d, permute=True, seed=11 + k, as_particles=True)
        for k, d in enumerate((mcm.Normal(0.0, 1e-12), mcm.Normal(np.pi / 3, 0.02), mcm.Normal(9.79, 0.02), mcm.Normal(1.0, 0.01), mcm.Normal(0.1, 0.01)))]
o3 = [mcm.particles._empty(n3, L.F64) for _ in range(2)]
trk = chains.pendulum_rk4_tape(0.01)
L.check(lib.mcm_chain_force_engine(0))
L.check(lib.mcm_chain_jit_option(1, 0)); L.check(lib.mcm_chain_jit_option(3, 1)); L.check(lib.mcm_chain_jit_option(4, 1))
for threads, ctas, packs in ((256, 2, 4), (320, 2, 4), (320, 2, 2), (384, 2, 2), (384, 1, 4), (192, 3, 4), (192, 3, 2), (192, 4, 2), (512, 1, 4), (128, 4, 4), (128, 5, 2), (448, 1, 4), (224, 2, 4), (288, 2, 4)):
    L.check(lib.mcm_chain_jit_option(5, threads)); L.check(lib.mcm_chain_jit_option(0, ctas)); L.check(lib.mcm_chain_jit_option(2, packs))
    try:
        ms, e = time_chain(t10, [a], [y], N, 0, t10.out_slots[0])
        ms2, e2 = time_chain(t10, [a], [y], N, 0, -1)
        print("generated threads=%d min_ctas=%d packs=%d : chain10+stats %.4f ms  chain10 no stats %.4f ms  (engines %d %d)" % (threads, ctas, packs, ms, ms2, e, e2), flush=True)
    except Exception as ex:
        print("generated threads=%d min_ctas=%d packs=%d : FAILED %s" % (threads, ctas, packs, ex), flush=True)
