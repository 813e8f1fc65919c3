"""Host twin of csrc/mcm_small_linalg_c.cuh — exp / log / det / eigvals / svdvals of small COMPLEX matrices and log of
real ones (complex.jl:104-181, particles.jl:603) — against LAPACK (numpy) and scipy on random batches: tolerance tier,
stated per operation.  CPU only; tests/test_gpu_linalg.py shows the device results equal the twin's bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "twin", "small_linalg_c_twin.cpp")
HDRS = [os.path.join(ROOT, "montecarlomeasurements.jl_b200", "csrc", f) for f in ("mcm_small_linalg_c.cuh", "mcm_small_linalg.cuh")]
SO = os.path.join(ROOT, "tests", "twin", "_build", "libslac_twin.so")
OPS = {"exp": 0, "log": 1, "det": 2, "eigvals": 3, "svdvals": 4}


def build_twin():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in [SRC] + HDRS):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO], check=True)
    return C.CDLL(SO)


def twin_run(lib, op, A):
    """A: (nb, n, n) complex or real -> complex results of the natural shape"""
    nb, n, _ = A.shape
    Ar = np.ascontiguousarray(A.real, dtype=np.float64)
    Ai = np.ascontiguousarray(A.imag, dtype=np.float64) if np.iscomplexobj(A) else None
    nout = {"exp": nb * n * n, "log": nb * n * n, "det": nb, "eigvals": nb * n, "svdvals": nb * n}[op]
    outr, outi = np.zeros(nout), np.zeros(nout)
    rc = lib.slac_twin(OPS[op], n, Ar.ctypes.data_as(C.c_void_p), Ai.ctypes.data_as(C.c_void_p) if Ai is not None else None,
                       outr.ctypes.data_as(C.c_void_p), outi.ctypes.data_as(C.c_void_p), C.c_long(nb))
    assert rc == 0, rc
    z = outr + 1j * outi
    if op in ("exp", "log"):
        return z.reshape(nb, n, n)
    if op == "det":
        return z
    return (outr if op == "svdvals" else z).reshape(nb, n)


def sort_spectrum(w):
    return np.array([sorted(r, key=lambda z: (round(z.real, 8), round(z.imag, 8))) for r in w])


@pytest.fixture(scope="module")
def twin():
    return build_twin()


@pytest.mark.parametrize("n", list(range(1, 9)) + [9, 12, 16])
def test_complex_twin_against_lapack(twin, n):
    rng = np.random.default_rng(100 + n)
    nb = 200
    A = rng.normal(size=(nb, n, n)) + 1j * rng.normal(size=(nb, n, n))
    d = twin_run(twin, "det", A)
    dr = np.linalg.det(A)
    assert np.max(np.abs(d - dr) / np.maximum(np.abs(dr), np.prod(np.linalg.norm(A, axis=2), axis=1) * 1e-3)) < 1e-12
    sv = twin_run(twin, "svdvals", A)
    sr = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(sv - sr) / sr[:, :1]) < 1e-13
    w = twin_run(twin, "eigvals", A)
    wr = np.linalg.eigvals(A)
    ws, rs = sort_spectrum(w), sort_spectrum(wr)
    # match spectra as multisets (ordering by rounded keys can differ for near-equal real parts)
    for a_, b_ in zip(w, wr):
        cost = np.abs(a_[:, None] - b_[None, :])
        assert np.max(np.min(cost, axis=1)) < 1e-10 * max(1.0, np.abs(b_).max()) and np.max(np.min(cost, axis=0)) < 1e-10 * max(1.0, np.abs(b_).max())
    assert np.all(np.diff(w.real, axis=1) >= 0)                                   # ordered by real part
    E = twin_run(twin, "exp", 0.7 * A)
    Er = np.array([sl.expm(0.7 * a_) for a_ in A])
    assert np.max(np.abs(E - Er) / np.max(np.abs(Er), axis=(1, 2), keepdims=True)) < 1e-12
    # log: principal logarithm of matrices with spectrum away from the negative real axis, and exp(log(A)) == A
    B = np.array([sl.expm(0.5 * a_) for a_ in A])
    Lg = twin_run(twin, "log", B)
    Lr = np.array([sl.logm(b_) for b_ in B])
    scale = np.max(np.abs(Lr), axis=(1, 2), keepdims=True) + 1.0
    cond = np.linalg.cond(B)[:, None, None]
    assert np.max(np.abs(Lg - Lr) / scale / cond) < 1e-12
    back = np.array([sl.expm(l_) for l_ in Lg])
    assert np.max(np.abs(back - B) / np.max(np.abs(B), axis=(1, 2), keepdims=True) / cond) < 1e-12


def test_log_of_real_matrices_like_the_reference(twin):
    """particles.jl:603 / runtests.jl: log of a real uncertain matrix is complex-valued in general; for a positive definite
    matrix it is real and exp(log(A)) == A; a negative eigenvalue gives Im = pi on the principal branch (Base.log)."""
    rng = np.random.default_rng(7)
    for n in (2, 3, 5):
        M = rng.normal(size=(100, n, n))
        S = M @ np.transpose(M, (0, 2, 1)) + 0.5 * np.eye(n)
        Lg = twin_run(twin, "log", S)
        Lr = np.array([sl.logm(s_) for s_ in S])
        assert np.max(np.abs(Lg.imag)) < 1e-12 and np.max(np.abs(Lg.real - Lr.real)) < 1e-11 * np.max(np.abs(Lr))
        G = twin_run(twin, "log", M)                                # general real matrices: complex logarithm
        back = np.array([sl.expm(g_) for g_ in G])
        cond = np.linalg.cond(M)[:, None, None]
        assert np.max(np.abs(back - M) / np.max(np.abs(M), axis=(1, 2), keepdims=True) / cond) < 1e-11
    D = np.diag([-2.0, 3.0])[None]
    Ld = twin_run(twin, "log", D)[0]
    assert abs(Ld[0, 0] - (np.log(2.0) + 1j * np.pi)) < 1e-14 and abs(Ld[1, 1] - np.log(3.0)) < 1e-14 and abs(Ld[0, 1]) < 1e-15
    Z = twin_run(twin, "log", np.zeros((1, 2, 2)))[0]
    assert not np.all(np.isfinite(Z))                               # singular: no logarithm
