"""Complex{Particles} on device (mcm_complex_op) against numpy's complex arithmetic on the same samples, plus the
reference's own assertions (test/runtests.jl:529-552).  + - * bit-exact (Base's formula, no contraction); division,
powers and functions within 1e-13 relative (tolerance tier)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(got, want):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_arithmetic_matches_base_formulas(mcm, dt):
    rng = np.random.default_rng(0)
    N = 100_003
    a = (rng.normal(size=N) + 1j * rng.normal(size=N))
    b = (rng.normal(size=N) + 1j * rng.normal(size=N))
    ar, ai, br, bi = (np.ascontiguousarray(x.astype(dt)) for x in (a.real, a.imag, b.real, b.imag))
    za = mcm.ComplexParticles(mcm.Particles(ar), mcm.Particles(ai))
    zb = mcm.ComplexParticles(mcm.Particles(br), mcm.Particles(bi))
    if dt is np.float64:
        s = (za + zb).to_numpy(); assert np.array_equal(s.real, ar + br) and np.array_equal(s.imag, ai + bi)
        d = (za - zb).to_numpy(); assert np.array_equal(d.real, ar - br) and np.array_equal(d.imag, ai - bi)
        m = (za * zb).to_numpy()
        assert np.array_equal(m.real, ar * br - ai * bi) and np.array_equal(m.imag, ar * bi + ai * br)
        m2 = (za * 2.5).to_numpy(); assert np.array_equal(m2.real, ar * 2.5 - ai * 0.0) and np.array_equal(m2.imag, ar * 0.0 + ai * 2.5)
        m3 = ((1 + 2j) * za).to_numpy(); assert np.array_equal(m3.real, 1.0 * ar - 2.0 * ai)
    tol = 1e-13 if dt is np.float64 else 1e-5
    A = ar.astype(np.float64) + 1j * ai.astype(np.float64)
    B = br.astype(np.float64) + 1j * bi.astype(np.float64)
    assert rel((za / zb).to_numpy(), A / B) < tol
    assert rel((za / 3.0).to_numpy(), A / 3.0) < tol
    assert rel((2.0 / zb).to_numpy(), 2.0 / B) < tol
    assert rel((mcm.Particles(ar) * zb).to_numpy(), ar.astype(np.float64) * B) < tol
    assert rel((-za).to_numpy(), -A) == 0 and rel(za.conj().to_numpy(), A.conj()) == 0
    assert rel(abs(za).to_numpy(), np.abs(A)) < tol and rel(za.angle().to_numpy(), np.angle(A)) < 10 * tol


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_functions_match_numpy(mcm, dt):
    rng = np.random.default_rng(1)
    N = 200_001
    A = (rng.normal(size=N) * 3 + 1j * rng.normal(size=N) * 3)
    A[:7] = [1.0, -1.0, 1e-8j, 0.9999999 + 1e-9j, -2.0 + 0.0j, 1j, 1.0000001]      # |z| ~ 1, the branch cut, pure real
    ar, ai = np.ascontiguousarray(A.real.astype(dt)), np.ascontiguousarray(A.imag.astype(dt))
    A = ar.astype(np.float64) + 1j * ai.astype(np.float64)
    z = mcm.ComplexParticles(mcm.Particles(ar), mcm.Particles(ai))
    C = mcm.complexp
    tol = 2e-13 if dt is np.float64 else 2e-5
    for f, ref in ((C.csqrt, np.sqrt), (C.cexp, np.exp), (C.clog, np.log), (C.csin, np.sin), (C.ccos, np.cos),
                   (C.ctan, np.tan), (C.clog10, np.log10), (C.cexp10, lambda x: np.exp(x * np.log(10.0)))):
        got, want = f(z).to_numpy(), ref(A)
        assert rel(got, want) < tol, f.__name__
    # log near |z| = 1: the real part is O(|z| - 1), absolute accuracy matters
    near = A[:7]
    assert np.max(np.abs(C.clog(z).to_numpy()[:7].real - np.log(near).real)) < (1e-15 if dt is np.float64 else 1e-6)
    # powers: complex^real, complex^complex, real^complex, integer exponent
    w = 0.7 - 0.3j
    assert rel((z ** 2.5).to_numpy(), A ** 2.5) < 10 * tol
    assert rel((z ** w).to_numpy(), A ** w) < 10 * tol
    assert rel((2.0 ** z).to_numpy(), 2.0 ** A) < 10 * tol
    assert rel((z ** 2).to_numpy(), A ** 2) < 10 * tol
    assert rel((z ** z).to_numpy()[7:2000], (A ** A)[7:2000]) < 1e3 * tol          # conditioning grows with |z log z|


def test_reference_assertions(mcm):
    """test/runtests.jl:529-552."""
    C = mcm.complexp
    p, q = mcm.pm(2.0, 0.0), mcm.pm(3.0, 0.0)
    zpp = mcm.complex_particles(p, p)
    assert np.allclose(C.csqrt(zpp).to_numpy(), np.sqrt(2 + 2j), rtol=1e-15, atol=0)
    assert np.allclose(C.cexp(zpp).to_numpy(), np.exp(2 + 2j), rtol=1e-15, atol=0)
    y = mcm.Particles(100)
    lhs = C.cexp(1j * mcm.complex_particles(y)).to_numpy()
    assert np.allclose(lhs, np.cos(y.to_numpy()) + 1j * np.sin(y.to_numpy()), rtol=1e-14, atol=1e-15)
    assert np.allclose((zpp / mcm.complex_particles(q, q)).to_numpy(), (2 + 2j) / (3 + 3j), rtol=1e-15)
    assert np.allclose((p / mcm.complex_particles(q, q)).to_numpy(), 2 / (3 + 3j), rtol=1e-15)
    assert np.allclose((2 / mcm.complex_particles(q, q)).to_numpy(), 2 / (3 + 3j), rtol=1e-15)
    assert not C.isinf(zpp) and C.isfinite(zpp)
    z = mcm.complex_particles(mcm.pm(1.0, 0.1), mcm.pm(1.0, 0.1))
    assert mcm.pmaximum(abs(C.csqrt(z ** 2) - z)) < 1e-15
    assert mcm.pmaximum(abs(C.csqrt(z ** 2.0) - z)) < 1e-15
    one = (z / z).to_numpy()
    assert np.allclose(one, 1.0, rtol=1e-15)
    z = mcm.complex_particles(mcm.pm(1.0, 0.1), mcm.pm(0.0, 0.1))
    assert np.allclose((2 ** z).re.to_numpy() if False else (2 ** z).to_numpy().real, (2.0 ** z.to_numpy()).real, rtol=1e-13)
    m = C.pmean(z)
    assert abs(m.real - 1.0) < 1e-12 and abs(m.imag) < 1e-12 and abs(C.pstd(z).real - mcm.pstd(z.re)) == 0


def test_errors(mcm):
    p = mcm.pm(1.0, 0.1)
    with pytest.raises(mcm.MethodError):
        mcm.complex_particles(p, mcm.Particles(10))
    z = mcm.complex_particles(p, p)
    with pytest.raises(mcm.MethodError):
        z + "a"
    L = mcm._lib
    assert L.lib().mcm_complex_op(999, L.F64, None, None, 0.0, 0.0, None, None, 0.0, 0.0, None, None, 0) == 0
    import ctypes as C
    o = mcm.Particles(4)
    assert L.lib().mcm_complex_op(999, L.F64, None, None, 0.0, 0.0, None, None, 0.0, 0.0, C.c_void_p(o.ptr), C.c_void_p(o.ptr), 4) == L.MCM_ERR_INVALID
