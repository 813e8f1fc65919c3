"""Probe: run the canonical chain10 tape through mcm_chain a few times (for ncu / timing)."""
import ctypes as C, importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mcm_lib", os.path.join(ROOT, "montecarlomeasurements.jl_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec); spec.loader.exec_module(L)
E = L.ENUMS; lib = L.init(0)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
stats_on = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dA = C.c_void_p(); dB = C.c_void_p()
L.check(lib.mcm_alloc(C.byref(dA), N * 8)); L.check(lib.mcm_alloc(C.byref(dB), N * 8))
L.check(lib.mcm_systematic_normal(L.F64, dA, N, np.pi, 0.1, None, 0, N))
U, B, K = E["MCM_TAPE_UNARY"], E["MCM_TAPE_BINARY"], E["MCM_TAPE_CONST"]
prog = [(E["MCM_TAPE_LOAD"], 0, 0, 0, 0), (U + E["MCM_U_SIN"], 1, 0, 0, 0), (U + E["MCM_U_SQR"], 2, 0, 0, 0),
        (B + E["MCM_B_MUL"], 1, 1, 2, 0), (U + E["MCM_U_NEG"], 2, 0, 0, 0), (U + E["MCM_U_EXP"], 2, 2, 0, 0),
        (B + E["MCM_B_ADD"], 1, 1, 2, 0), (U + E["MCM_U_ABS"], 1, 1, 0, 0), (U + E["MCM_U_SQRT"], 1, 1, 0, 0),
        (B + E["MCM_B_POW"], 2, 0, K | 0, 0), (B + E["MCM_B_DIV"], 1, 1, 2, 0), (E["MCM_TAPE_STORE"], 0, 1, 0, 0)]
tape = (L.TapeIns * len(prog))()
for i, (op, d, a, b, c) in enumerate(prog):
    tape[i].op, tape[i].dst, tape[i].a, tape[i].b, tape[i].c = op, d, a, b, c
consts = np.array([1.5]); stats = np.zeros(6)
ins = (C.c_void_p * 1)(dA); outs = (C.c_void_p * 1)(dB)
for r in range(reps):
    t = time.perf_counter()
    L.check(lib.mcm_chain(L.F64, tape, len(prog), len(prog), len(prog), 0, consts.ctypes.data_as(L._dp), 1, ins, 1, outs, 1, N,
                          1 if stats_on else -1, stats.ctypes.data_as(L._dp) if stats_on else None))
    L.check(lib.mcm_sync())
    dt = time.perf_counter() - t
    print(f"chain10 N={N} stats={stats_on}: {dt*1e3:.3f} ms {10*N/dt/1e9:.1f} Gpop/s", stats[:3])
