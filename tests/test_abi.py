"""The C-ABI boundary: libmcm_b200.so loads, exports every function include/mcm_b200.h declares, the
Python binding covers exactly that list, and the product path fails LOUDLY (no CPU fallback) when no GPU
is present.  No compute calls — runs on the CPU-only box."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcm_b200.h")
PKG = os.path.join(ROOT, "montecarlomeasurements.jl_b200")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    txt = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", txt, flags=re.S)
    txt = re.sub(r"enum\s*\w*\s*\{.*?\}\s*;", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcm_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def L():
    import mcm_b200
    return mcm_b200._lib


def test_library_is_built_in_tree(L):
    assert os.path.exists(L.LIB_PATH), "run python montecarlomeasurements.jl_b200/build.py"
    assert os.path.dirname(L.LIB_PATH) == PKG


def test_every_declared_symbol_is_exported_and_bound(L):
    decl = declared_functions()
    assert len(decl) >= 50
    lib = L.load()
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/mcm_b200.h but not exported"
    assert sorted(L.SIGNATURES) == decl, set(L.SIGNATURES) ^ set(decl)
    assert lib.mcm_version() == L.ENUMS["MCM_B200_VERSION"]


def test_built_for_sm100a_only(L):
    out = subprocess.run(["cuobjdump", "--list-elf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_enums_match_header(L):
    E = L.ENUMS
    assert E["MCM_U_PLUS"] == 0 and E["MCM_U_FLOAT"] == 34 and E["MCM_U_CUBE"] == 41     # 35 registered + 7 extra
    assert E["MCM_U_COUNT"] == 49 and E["MCM_U_COUNT"] <= 0x40                            # + 3 round modes + 4 rem2pi; tape opcode space
    assert E["MCM_B_FDIV"] == 15 and E["MCM_B_COUNT"] == 24 and E["MCM_B_COUNT"] <= 0x20 and E["MCM_T_COUNT"] == 3
    assert E["MCM_FLAG_BOUNDS"] == 0x80000000 and E["MCM_CMP_ISNAN"] == 8 and E["MCM_RM_FROMZERO"] == 6
    assert E["MCM_TAPE_CONST"] == 0x40 and E["MCM_TAPE_SLOTS"] == 16
    assert C.sizeof(L.TapeIns) == 8


def test_no_cpu_fallback_without_gpu(L):
    """On a box without a CUDA device the product path must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    rc = lib.mcm_init(0)
    assert rc == L.MCM_ERR_CUDA
    assert b"no CPU fallback" in lib.mcm_last_error()
    x = np.zeros(4)
    assert lib.mcm_unary(0, L.F64, x.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 4) == L.MCM_ERR_NOT_INIT
    with pytest.raises(L.MCMError):
        L.init(0)
    import mcm_b200
    with pytest.raises(L.MCMError):
        mcm_b200.Particles(np.arange(4.0))


def test_moments_merge_host_helper(L):
    """mcm_moments_merge is pure host code: Chan merge must agree with the Python mirror used across ranks."""
    from mcm_b200 import sharded
    lib = L.load()
    rng = np.random.default_rng(0)
    x = rng.normal(3, 2, 1000); a, b = x[:300], x[300:]

    def mom(v):
        return np.array([v.size, v.mean(), ((v - v.mean()) ** 2).sum(), v.min(), v.max(), v.sum()])
    out = np.zeros(6)
    ma, mb = mom(a), mom(b)
    assert lib.mcm_moments_merge(ma.ctypes.data_as(L._dp), mb.ctypes.data_as(L._dp), out.ctypes.data_as(L._dp)) == 0
    assert np.allclose(out, mom(x), rtol=1e-13)
    assert np.array_equal(out, sharded.merge_moments(ma, mb))
    empty = np.array([0, 0, 0, np.inf, -np.inf, 0.0])
    assert lib.mcm_moments_merge(empty.ctypes.data_as(L._dp), mb.ctypes.data_as(L._dp), out.ctypes.data_as(L._dp)) == 0
    assert np.array_equal(out, mb)
