"""Time the Float64 elementwise primitives at N = 1e8 (kernel-only, CUDA events): python scripts/ab_unary.py"""
import ctypes as C
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mcm_b200 as mcm
L = mcm._lib
mcm.init(0)
lib = L.lib()
stream = torch.cuda.Stream()
L.check(lib.mcm_set_stream(C.c_void_p(stream.cuda_stream)))
N = 100_000_000
a = mcm.systematic_sample(N, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)
w = mcm.systematic_sample(N, mcm.Uniform(-50.0, 50.0), permute=True, seed=1, as_particles=True)
y = mcm.particles._empty(N, L.F64)
for name, op, src in (("sin", "SIN", a), ("sin(wide)", "SIN", w), ("cos", "COS", a), ("exp", "EXP", a), ("sqrt", "SQRT", a), ("inv", "INV", a), ("log", "LOG", a)):
    code = L.ENUMS["MCM_U_" + op]
    with torch.cuda.stream(stream):
        for _ in range(3):
            L.check(lib.mcm_unary(code, L.F64, C.c_void_p(src.ptr), C.c_void_p(y.ptr), N))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            L.check(lib.mcm_unary(code, L.F64, C.c_void_p(src.ptr), C.c_void_p(y.ptr), N))
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("MCM_MAPFAST=%s %-10s %.4f ms  %.0f GB/s" % (os.environ.get("MCM_MAPFAST", "0"), name, ms, 16 * N / ms / 1e6), flush=True)
