"""Host-side logic that needs no device: the expression tracer / slot allocator, the ahead-of-time chain
registry, shard ranges, constructor promotion rules and error classes."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def m():
    import mcm_b200
    return mcm_b200


def run_tape_numpy(t, inputs, E, reps=0):
    """Tiny numpy interpreter of the tape ABI (test-side only) to check the tracer's output."""
    import math
    U, B, T_ = E["MCM_TAPE_UNARY"], E["MCM_TAPE_BINARY"], E["MCM_TAPE_TERNARY"]
    un = {E["MCM_U_SIN"]: np.sin, E["MCM_U_EXP"]: np.exp, E["MCM_U_NEG"]: np.negative, E["MCM_U_SQR"]: np.square,
          E["MCM_U_ABS"]: np.abs, E["MCM_U_SQRT"]: np.sqrt, E["MCM_U_COS"]: np.cos, E["MCM_U_PLUS"]: lambda x: x,
          E["MCM_U_CUBE"]: lambda x: x * x * x, E["MCM_U_INV"]: lambda x: 1 / x}
    bi = {E["MCM_B_ADD"]: np.add, E["MCM_B_SUB"]: np.subtract, E["MCM_B_MUL"]: np.multiply, E["MCM_B_DIV"]: np.divide,
          E["MCM_B_POW"]: np.power}
    slots = {}; outs = {}

    def val(o):
        return np.full_like(inputs[0], t.consts[o & 31]) if o & E["MCM_TAPE_CONST"] else slots[o]

    def run(lo, hi):
        for (op, d, a, b, c) in t.ins[lo:hi]:
            if op == E["MCM_TAPE_LOAD"]: slots[d] = inputs[a].copy()
            elif op == E["MCM_TAPE_STORE"]: outs[d] = val(a).copy()
            elif op == E["MCM_TAPE_MOV"]: slots[d] = val(a).copy()
            elif op == E["MCM_TAPE_POWI"]: slots[d] = val(a) ** (b - 256 if b > 127 else b)
            elif op < B: slots[d] = un[op - U](val(a))
            elif op < T_: slots[d] = bi[op - B](val(a), val(b))
            else: raise AssertionError(op)
    lb, le = t.bounds
    run(0, lb)
    for _ in range(reps): run(lb, le)
    run(le, len(t.ins))
    return [outs[k] for k in sorted(outs)]


def test_tracer_chain10_semantics_and_slots(m):
    from mcm_b200 import chains, tape
    E = m._lib.ENUMS
    t = tape.trace(chains.chain10, 1)
    assert t.nin == 1 and t.nout == 1 and t.consts == [1.5]
    assert len(t.ins) == 12                          # LOAD + 10 primitives + STORE
    assert max(i[1] for i in t.ins if i[0] != E["MCM_TAPE_STORE"]) <= 2   # liveness: three slots suffice
    a = np.pi + 0.1 * np.random.default_rng(0).normal(size=100)
    (y,) = run_tape_numpy(t, [a], E)
    assert np.allclose(y, np.sqrt(np.abs(np.sin(a) * a ** 2 + np.exp(-a))) / a ** 1.5, rtol=1e-14)


def test_tracer_constants_scalars_and_multiple_outputs(m):
    from mcm_b200 import tape
    E = m._lib.ENUMS
    t = tape.trace(lambda a, b: (2.0 * a + b / 4.0 - 1.0, (a - b) ** 3, a ** -2), 2)
    assert t.nout == 3 and sorted(t.consts) == [-1.0, 2.0, 4.0] or len(t.consts) == 3
    a = np.linspace(1, 2, 7); b = np.linspace(-1, 1, 7)
    y0, y1, y2 = run_tape_numpy(t, [a, b], E)
    assert np.allclose(y0, 2 * a + b / 4 - 1) and np.allclose(y1, (a - b) ** 3) and np.allclose(y2, a ** -2.0)
    with pytest.raises(m.ArgumentError):
        tape.trace(lambda a: sum((a * float(k)) for k in range(40)), 1)   # > 32 distinct constants


def test_tracer_runs_out_of_slots_cleanly(m):
    from mcm_b200 import tape

    def wide(a):
        vals = [m.sin(a * float(k + 1)) for k in range(20)]   # 20 simultaneously live values
        out = vals[0]
        for v in vals[1:]:
            out = out * v
        return out
    with pytest.raises(m.ArgumentError):
        tape.trace(wide, 1)


def test_loop_tape_pendulum_matches_host_rk4(m, orc):
    from mcm_b200 import chains
    E = m._lib.ENUMS
    t = chains.pendulum_rk4_tape(0.01)
    lb, le = t.bounds
    assert 0 < lb < le < len(t.ins) and t.nin == 5 and t.nout == 2
    assert max(i[1] for i in t.ins if i[0] != E["MCM_TAPE_STORE"]) < E["MCM_TAPE_SLOTS"]
    rng = np.random.default_rng(1)
    n = 50
    th = np.zeros(n); om = np.pi / 3 + 0.02 * rng.normal(size=n)
    g = 9.79 + 0.02 * rng.normal(size=n); Lp = 1 + 0.01 * rng.normal(size=n); c = 0.1 + 0.01 * rng.normal(size=n)
    th1, om1 = run_tape_numpy(t, [th, om, g, Lp, c], E, reps=25)
    th2, om2 = orc.pendulum_rk4(th, om, g, Lp, c, 0.01, 25)
    assert np.allclose(th1, th2, rtol=1e-13, atol=1e-15) and np.allclose(om1, om2, rtol=1e-13)


@pytest.mark.parametrize("name,step,ref", [
    ("fib", lambda a, b: (a + b, a), lambda a, b: (a + b, a)),
    ("swap", lambda a, b: (b, a), lambda a, b: (b, a)),
    ("rot3", lambda a, b, c: (b, c, a), lambda a, b, c: (b, c, a)),
    ("mixed", lambda a, b, c: (c, a * b, a), lambda a, b, c: (c, a * b, a)),
    ("keep", lambda a, b: (a, a + b), lambda a, b: (a, a + b)),
])
def test_loop_writeback_is_a_parallel_assignment(m, name, step, ref):
    """(a, b) -> (a + b, a) and pure permutations read OLD state values (tape.py trace_loop)."""
    from mcm_b200 import tape
    E = m._lib.ENUMS
    ns = step.__code__.co_argcount
    t = tape.trace_loop(step, ns, 0)
    rng = np.random.default_rng(5)
    x0 = [rng.normal(size=9) for _ in range(ns)]
    got = run_tape_numpy(t, [v.copy() for v in x0], E, reps=5)
    want = [v.copy() for v in x0]
    for _ in range(5):
        want = list(ref(*want))
    for g, w in zip(got, want):
        assert np.array_equal(g, w), name


def test_chain_registry_is_up_to_date():
    """csrc/chain_registry.def must be what scripts/gen_chain_registry.py generates from chains.py."""
    path = os.path.join(ROOT, "montecarlomeasurements.jl_b200", "csrc", "chain_registry.def")
    before = open(path).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_chain_registry.py")], check=True, capture_output=True)
    assert open(path).read() == before, "regenerate the registry and rebuild"


def test_shard_ranges_cover_and_align(m):
    for N in (1, 2, 7, 100, 100_000_000, 10_000_001):
        for g in (1, 2, 4, 8):
            spans = [m.sharded.shard_range(N, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % 2 == 0 for lo, hi in spans if hi > lo)  # 16-byte aligned Float64 shard starts
            sizes = [hi - lo for lo, hi in spans if hi > lo]
            assert max(sizes) - min(sizes) <= 2 * g or len(sizes) < g
    with pytest.raises(m.ArgumentError):
        m.sharded.shard_range(10, 3, 2)


def test_distribution_promotion_rules(m):
    # ±: T = promote_type(float(typeof μ), float(typeof σ)) (particles.jl:20)
    assert m.Normal(np.float32(1), np.float32(2)).eltype == np.float32
    assert m.Normal(np.float32(1), 2.0).eltype == np.float64
    assert m.Normal(1, 2).eltype == np.float64
    mv = m.MvNormal([1.0, 2.0], 0.5)
    assert np.allclose(mv.S, 0.25 * np.eye(2))                     # scalar σ is a std
    assert np.allclose(m.MvNormal([1.0, 2.0], [4.0, 9.0]).S, np.diag([4.0, 9.0]))  # vector = variances
    assert m.DEFAULT_NUM_PARTICLES == 2000 and m.DEFAULT_STATIC_NUM_PARTICLES == 100


def test_error_classes(m):
    assert issubclass(m.ArgumentError, ValueError) and issubclass(m.MethodError, TypeError)
    assert issubclass(m.ArgumentError, m.MCMError)


def test_chain_kernel_generation_and_nvrtc_compile_without_a_gpu(m, tmp_path):
    """csrc/chain_jit.cu: the tape -> CUDA C++ generator and its NVRTC compilation need no device."""
    import ctypes as C
    import subprocess
    from mcm_b200 import chains, tape
    L = m._lib
    lib = L.load()
    for name, t, f64 in (("chain10", tape.trace(chains.chain10, 1), True), ("rk4", chains.pendulum_rk4_tape(0.01), True),
                         ("chain10_f32", tape.trace(chains.chain10, 1), False)):
        src = C.create_string_buffer(1 << 18)
        nb = C.c_size_t(0)
        consts = np.asarray(t.consts if t.consts else [0.0])
        lb, le = t.bounds
        cub = str(tmp_path / (name + ".cubin"))
        rc = lib.mcm_chain_jit_source(L.F64 if f64 else L.F32, t.as_ctypes(), len(t.ins), lb, le, consts.ctypes.data_as(L._dp),
                                      len(t.consts), t.out_slots[0], 1, src, len(src), C.byref(nb), cub.encode())
        assert rc == 0, lib.mcm_last_error()
        text = src.value.decode()
        assert "mcm_jit_kernel" in text and nb.value > 10000
        if name == "chain10":
            assert "pow_three_halves_fast" in text and "sincos_fast_v" in text and "exp_fast" in text   # structural constant folded
            assert "sqrt_fast" in text and "div_fast" in text and "ok = ok &&" in text   # inline sqrt / division, window predicate
        if name == "chain10_f32":
            assert "pow_three_halves" not in text                                              # Float32 keeps the library pow
        if name == "rk4":
            assert "for (int rep = 0; rep < p.reps; ++rep)" in text
        out = subprocess.run(["cuobjdump", "--list-elf", cub], capture_output=True, text=True).stdout
        assert "sm_100a" in out
    # an invalid tape is rejected by the same validator as mcm_chain
    bad = (L.TapeIns * 1)()
    bad[0].op = 0x7f
    assert lib.mcm_chain_jit_source(L.F64, bad, 1, 1, 1, None, 0, -1, 0, None, 0, None, None) != 0


@pytest.mark.parametrize("n", [100000, 1 << 18, 300007])
def test_index_bijection_statistics(orc, n):
    """The keyed bijection that stands in for randperm (sampling.jl:11): columns are permutations, pairwise rank
    correlations / correlation with the identity / between neighbours stay inside 4.5 / sqrt(n), and the 32 x 32 joint
    position tables (column vs column, column vs identity) are uniform (chi-square within 6 sigma)."""
    cols = [orc.index_bijection(n, 5, c) for c in range(5)]
    ident = np.arange(n)
    for p in cols:
        assert np.array_equal(np.sort(p), ident)
    lim = 4.5 / np.sqrt(n)
    for a in range(5):
        assert abs(np.corrcoef(cols[a], ident)[0, 1]) < lim
        assert abs(np.corrcoef(cols[a][:-1], cols[a][1:])[0, 1]) < lim
        for b in range(a):
            assert abs(np.corrcoef(cols[a], cols[b])[0, 1]) < lim
    B = 32
    dof = (B - 1) * (B - 1)

    def chi2(p, q):
        t = np.bincount((p * B // n) * B + (q * B // n), minlength=B * B).astype(float)
        rp = np.bincount(p * B // n, minlength=B).astype(float)
        e = np.outer(rp, rp) / n
        return ((t.reshape(B, B) - e) ** 2 / e).sum()
    for a in range(3):
        assert abs(chi2(cols[a], ident) - dof) < 6 * np.sqrt(2 * dof)
        for b in range(a):
            assert abs(chi2(cols[a], cols[b]) - dof) < 6 * np.sqrt(2 * dof)
