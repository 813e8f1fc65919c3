"""Host twin of the hand-written Float64 device math (csrc/mcm_fastmath.cuh, csrc/mcm_norminv.cuh): the
same source compiled as plain C++ with contraction off is bit-identical to the sm_100a code (the library is
built with -fmad=false and spells every FMA), so ulp bounds are measured here on millions of points and the
GPU test only has to show bit-equality with the twin.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ulp_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "twin", "fastmath_twin.cpp")
SO = os.path.join(ROOT, "tests", "twin", "_build", "libfastmath_twin.so")


def build_twin():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "montecarlomeasurements.jl_b200", "csrc", f)
                    for f in ("mcm_fastmath.cuh", "mcm_fastmath_coeffs.inc", "mcm_norminv.cuh", "mcm_norminv_coeffs.inc")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO], check=True)
    lib = C.CDLL(SO)
    lib.twin_maxulp.restype = C.c_double
    return lib


@pytest.fixture(scope="module")
def twin():
    return build_twin()


def call1(lib, name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    getattr(lib, name)(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_long(x.size))
    return y


def maxulp(lib, fn, x, e=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    wx = C.c_double()
    ep = np.ascontiguousarray(e, dtype=np.float64).ctypes.data_as(C.c_void_p) if e is not None else None
    return lib.twin_maxulp(fn, x.ctypes.data_as(C.c_void_p), ep, C.c_long(x.size), C.byref(wx)), wx.value


def test_sincos_kernel_forms_return_the_same_bits(twin):
    """Generated kernels pick the sin / cos kernel per warp (fixed parity or selected operands): the result of a lane
    must not depend on its neighbours, so all forms must agree bit for bit."""
    rng = np.random.default_rng(3)
    n = 2_000_000
    r = rng.uniform(-0.7854, 0.7854, n); r[:1000] = rng.normal(0, 1e-9, 1000); r[1000] = 0.0; r[1001] = -0.0
    rl = r * rng.uniform(-1, 1, n) * 2.0 ** -53
    twin.twin_sincos_forms_differ.restype = C.c_long
    assert twin.twin_sincos_forms_differ(r.ctypes.data_as(C.c_void_p), rl.ctypes.data_as(C.c_void_p), C.c_long(n)) == 0


def test_sin_cos_fast_range(twin):
    rng = np.random.default_rng(0)
    n = 1_000_000
    k = rng.integers(1, 130000, n).astype(np.float64)
    near = k * (np.pi / 2); near = near + np.spacing(near) * rng.integers(-3, 4, n)   # doubles hugging k*pi/2
    for fn in (0, 1):
        for x in (rng.uniform(-8, 8, n), np.pi + rng.normal(0, 0.1, n), rng.uniform(-2e5, 2e5, n), near,
                  rng.normal(0, 1e-8, n)):
            u, wx = maxulp(twin, fn, x)
            assert u <= 0.80, (fn, u, wx)
    s = call1(twin, "twin_sin", np.array([0.0, -0.0, np.inf, np.nan, 1e300, 3e5]))
    assert s[0] == 0 and not np.signbit(s[0]) and np.signbit(s[1]) and np.isnan(s[2]) and np.isnan(s[3])
    assert abs(s[4] - np.sin(1e300)) <= 2 * np.spacing(abs(np.sin(1e300))) and abs(s[5] - np.sin(3e5)) < 1e-15  # library path


def test_exp(twin):
    rng = np.random.default_rng(1)
    n = 1_000_000
    for x in (rng.uniform(-5, 5, n), rng.uniform(-708, 708, n), -np.pi + rng.normal(0, 0.1, n), rng.normal(0, 1e-9, n)):
        u, wx = maxulp(twin, 2, x)
        assert u <= 0.85, (u, wx)
    e = call1(twin, "twin_exp", np.array([0.0, 710.0, -745.0, -760.0, np.inf, -np.inf, np.nan, 709.7, -708.3]))
    assert e[0] == 1.0 and np.isinf(e[1]) and e[2] > 0 and e[3] == 0 and np.isinf(e[4]) and e[5] == 0 and np.isnan(e[6])
    assert ulp_err(e[7:], np.exp(np.array([709.7, -708.3]))).max() <= 1


def test_pow_half_integer_exponents_are_correctly_rounded(twin):
    rng = np.random.default_rng(2)
    n = 500_000
    for y in (1.5, 0.5, 2.5, 3.5, 7.5, -0.5, -1.5, -2.5, 8.5):
        for x in (np.abs(np.pi + rng.normal(0, 0.3, n)), np.exp(rng.uniform(-60, 60, n))):
            u, wx = maxulp(twin, 3, x, np.full(n, y))
            assert u <= 0.502, (y, u, wx)
    # everything else goes to the library pow with identical special cases
    x = np.array([2.0, -2.0, 0.0, np.inf, 4.0, 1e-300, 1e300, np.nan])
    y = np.array([1.5, 1.5, 1.5, 1.5, 0.3, 1.5, 1.5, 1.5])
    out = np.empty(8)
    twin.twin_pow(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(8))
    assert out[0] == 2.0 ** 1.5 and np.isnan(out[1]) and out[2] == 0 and np.isinf(out[3]) and np.isnan(out[7])
    assert out[4] == 4.0 ** 0.3 and out[5] == (1e-300) ** 1.5 and np.isinf(out[6])


def test_against_mpmath_golden(twin, golden):
    for name, fn, lim in (("sin", "twin_sin", 1), ("cos", "twin_cos", 1), ("exp", "twin_exp", 1)):
        u = ulp_err(call1(twin, fn, golden[name + "_x"]), golden[name + "_y"])
        assert u.max() <= lim, (name, u.max())
    half = golden["pow_e"] == 1.5
    x = np.ascontiguousarray(golden["pow_x"][half]); e = np.full(x.size, 1.5); out = np.empty_like(x)
    twin.twin_pow(x.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(x.size))
    assert np.array_equal(out, golden["pow_y"][half])                  # a^1.5: correctly rounded on the fixture
    u = ulp_err(call1(twin, "twin_norminv", golden["norminv_p"]), golden["norminv_y"])
    assert u.max() <= 1, u.max()


def test_norminv_below_one_ulp_everywhere(twin):
    """quantile(Normal(0,1), p) (src/sampling.jl:10) against the double-double truth of
    tests/golden/norminv_truth.npz: grids, tails down to the subnormals, p next to 1/2 and 1, every branch and
    piece boundary of csrc/mcm_norminv.cuh.  north_star tolerance is 2 ulp; AS 241 measured 5."""
    t = np.load(os.path.join(ROOT, "tests", "golden", "norminv_truth.npz"))
    got = call1(twin, "twin_norminv", t["p"])
    u = np.abs((got - t["y"]) - t["y_lo"]) / np.spacing(np.abs(t["y"]))
    assert u.max() <= 0.75, (u.max(), t["p"][np.argmax(u)])
    s = call1(twin, "twin_norminv", np.array([0.0, 1.0, -0.1, 1.1, np.nan, 0.5]))
    assert np.isneginf(s[0]) and np.isposinf(s[1]) and np.isnan(s[2:5]).all() and s[5] == 0.0
    # monotone on a fine grid across all the piece boundaries
    p = np.exp(-np.linspace(1.9, 48.0, 200001) ** 2 / 2)
    z = call1(twin, "twin_norminv", p[p > 0])
    assert np.all(np.diff(z) <= 0)


def test_rem2pi_is_correctly_rounded(twin):
    """rem2pi(x, r) (src/particles.jl:359) for all four rounding modes against exact arithmetic at 1600 bits."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    rng = np.random.default_rng(3)
    k = rng.integers(1, 10 ** 6, 300).astype(float)
    near = k * (np.pi / 2); near = near + np.spacing(near) * rng.integers(-2, 3, 300)
    x = np.concatenate([rng.normal(0, 3, 500), rng.uniform(-2e5, 2e5, 500), 10.0 ** rng.uniform(5, 300, 500) * rng.choice([-1, 1], 500),
                        near, -near, [0.0, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 1e22, -1e22, 1.7e308, 205887.0, 205888.0]])
    for mode, name in enumerate(("RoundNearest", "RoundToZero", "RoundDown", "RoundUp")):
        y = np.empty_like(x)
        twin.twin_rem2pi(x.ctypes.data_as(C.c_void_p), C.c_int(mode), y.ctypes.data_as(C.c_void_p), C.c_long(x.size))
        ref = orc.rem2pi_truth(x, name)
        assert ulp_err(y, ref).max() <= 1 and (y != ref).mean() < 1e-3, name
