"""mcm_sort / mcm_sortperm: the fused form (first pass reads the particles, last pass writes values and Int64 indices)
against numpy and against the separate-pass form, bit for bit — Julia `isless` order (-0.0 < +0.0, NaN last;
src/particles.jl:316-323 forwards to Base sort / Statistics.quantile), stable permutation."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def isless_key(x):
    """Order-preserving integer key of a float array (the oracle's definition: sign flip, NaN = all ones)."""
    u = x.view(np.uint64 if x.dtype == np.float64 else np.uint32)
    top = u.dtype.type(1) << u.dtype.type(8 * u.itemsize - 1)
    k = np.where(u & top, ~u, u | top)
    k[np.isnan(x)] = ~u.dtype.type(0)
    return k


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 5, 4095, 4096, 4097, 100003, (1 << 20) + 7, 3_000_017, 5 * 296 * 4096])  # the last two: several tiles per persistent CTA
def test_fused_sort_matches_numpy_and_the_separate_passes(mcm, dt, n):
    lib = mcm._lib.lib()
    rng = np.random.default_rng(n)
    x = rng.normal(0.0, 3.0, n).astype(dt)
    if n > 100:
        x[rng.integers(0, n, 20)] = np.nan
        x[rng.integers(0, n, 20)] = 0.0
        x[rng.integers(0, n, 20)] = -0.0
        x[rng.integers(0, n, 5)] = np.inf
        x[rng.integers(0, n, 5)] = -np.inf
        x[rng.integers(0, n, 50)] = x[0]            # ties: the permutation must be stable
    p = mcm.Particles(x)
    order = np.argsort(isless_key(x), kind="stable")
    want = x[order]
    res = {}
    for fused in (1, 0):
        lib.mcm_sort_use_fused(fused)
        s = mcm.psort(p).to_numpy()
        perm = mcm.psortperm(p).to_numpy()
        res[fused] = (s, perm)
        assert np.array_equal(isless_key(s), isless_key(want)), f"fused={fused}"
        assert np.array_equal(perm, order), f"fused={fused}"
    lib.mcm_sort_use_fused(1)
    assert np.array_equal(res[0][0].view(np.uint8), res[1][0].view(np.uint8))
    try:
        mcm.check_domain()            # NaN keys raise the sticky quantile flag: read and reset it
    except Exception:
        pass


def test_constant_low_digit_falls_back(mcm):
    """Integer-valued doubles share their low mantissa byte: the fused first pass has nothing to sort on and the
    separate-pass form runs — same result."""
    rng = np.random.default_rng(1)
    x = rng.integers(-1000, 1000, 50000).astype(np.float64)
    p = mcm.Particles(x)
    assert np.array_equal(mcm.psort(p).to_numpy(), np.sort(x))
    assert np.array_equal(mcm.psortperm(p).to_numpy(), np.argsort(x, kind="stable"))
    c = mcm.Particles(np.full(10000, 2.5))                   # every digit constant: no pass at all
    assert np.array_equal(mcm.psort(c).to_numpy(), np.full(10000, 2.5))
    assert np.array_equal(mcm.psortperm(c).to_numpy(), np.arange(10000))


@pytest.mark.parametrize("kind", ["equal", "two", "sorted", "reversed", "nan_heavy", "tiny_range"])
def test_adversarial_key_distributions(mcm, kind):
    """Digit histograms that are as skewed as they get (one bin, two bins, long runs): every tile of a pass lands in the
    same few buckets.  n spans several tiles per persistent CTA."""
    n = 2_500_003
    rng = np.random.default_rng(3)
    x = {"equal": np.full(n, -1.25), "two": rng.choice([1.0, -2.0], n), "sorted": np.sort(rng.normal(size=n)),
         "reversed": np.sort(rng.normal(size=n))[::-1].copy(), "nan_heavy": np.where(rng.random(n) < 0.5, np.nan, rng.normal(size=n)),
         "tiny_range": 1.0 + rng.integers(0, 3, n) * 2.0 ** -52}[kind]
    p = mcm.Particles(np.ascontiguousarray(x))
    order = np.argsort(isless_key(x), kind="stable")
    assert np.array_equal(isless_key(mcm.psort(p).to_numpy()), isless_key(x[order]))
    assert np.array_equal(mcm.psortperm(p).to_numpy(), order)
    try:
        mcm.check_domain()
    except Exception:
        pass
