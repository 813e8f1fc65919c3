"""Host twin of csrc/mcm_small_linalg.cuh (the per-thread algorithms behind mcm_batched_linalg) against LAPACK
(numpy) and scipy on random batches: tolerance tier, stated per operation.  CPU only; the GPU test
(tests/test_gpu_linalg.py) then shows the device results equal the twin's bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "twin", "small_linalg_twin.cpp")
HDR = os.path.join(ROOT, "montecarlomeasurements.jl_b200", "csrc", "mcm_small_linalg.cuh")
SO = os.path.join(ROOT, "tests", "twin", "_build", "libsla_twin.so")
OPS = {"solve": 0, "inv": 1, "det": 2, "exp": 3, "eigvals": 4, "eigvals_sym": 5, "lyap": 6, "svdvals": 7}


def build_twin():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in (SRC, HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO], check=True)
    return C.CDLL(SO)


def twin_run(lib, op, A, B=None, m=0, nout=0, im=False):
    nb, n, _ = A.shape
    A = np.ascontiguousarray(A, dtype=np.float64)
    out, oi = np.zeros(nout), np.zeros(nout)
    Bp = np.ascontiguousarray(B, dtype=np.float64).ctypes.data_as(C.c_void_p) if B is not None else None
    rc = lib.sla_twin(OPS[op], n, A.ctypes.data_as(C.c_void_p), Bp, m, out.ctypes.data_as(C.c_void_p),
                      oi.ctypes.data_as(C.c_void_p), C.c_long(nb))
    assert rc == 0, rc
    return (out, oi) if im else out


def sort_spectrum(w):
    return np.array([sorted(r, key=lambda z: (round(z.real, 9), z.imag)) for r in w])


@pytest.fixture(scope="module")
def twin():
    return build_twin()


@pytest.mark.parametrize("n", list(range(1, 9)) + [9, 12, 16])
def test_twin_against_lapack(twin, n):
    rng = np.random.default_rng(n)
    nb = 500
    A = rng.normal(size=(nb, n, n))
    B = rng.normal(size=(nb, n, 3))
    cond = np.linalg.cond(A)[:, None, None]
    X = twin_run(twin, "solve", A, B, 3, nb * n * 3).reshape(nb, n, 3)
    ref = np.linalg.solve(A, B)
    assert np.max(np.abs(X - ref) / np.max(np.abs(ref), axis=(1, 2), keepdims=True) / cond) < 1e-14
    Xi = twin_run(twin, "inv", A, nout=nb * n * n).reshape(nb, n, n)
    ri = np.linalg.inv(A)
    assert np.max(np.abs(Xi - ri) / np.max(np.abs(ri), axis=(1, 2), keepdims=True) / cond) < 1e-14
    d = twin_run(twin, "det", A, nout=nb)
    dr = np.linalg.det(A)
    assert np.max(np.abs(d - dr) / np.abs(dr)) < 1e-10
    As = A * rng.choice([0.01, 0.3, 1.0, 3.0, 20.0], size=(nb, 1, 1))      # every Padé order and the squaring branch
    Ex = twin_run(twin, "exp", As, nout=nb * n * n).reshape(nb, n, n)
    Er = np.array([sl.expm(a) for a in As])
    assert np.max(np.abs(Ex - Er) / np.max(np.abs(Er), axis=(1, 2), keepdims=True)) < 1e-9     # the reference's own bar
    wr, wi = twin_run(twin, "eigvals", A, nout=nb * n, im=True)
    w = sort_spectrum((wr + 1j * wi).reshape(nb, n))
    ref = sort_spectrum(np.linalg.eigvals(A))
    assert np.max(np.abs(w - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) < 1e-11
    S = A + A.transpose(0, 2, 1)
    ws = twin_run(twin, "eigvals_sym", S, nout=nb * n).reshape(nb, n)
    rs = np.linalg.eigvalsh(S)
    assert np.max(np.abs(ws - rs) / np.max(np.abs(rs), axis=1, keepdims=True)) < 1e-13
    sv = twin_run(twin, "svdvals", A, nout=nb * n).reshape(nb, n)
    svr = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(sv - svr) / svr[:, :1]) < 1e-13
    if n <= 8:
        Ast = A - max(3.0, 1.5 * np.sqrt(n) + 1.5) * np.eye(n)      # keeps lambda_i + lambda_j away from 0 as n grows
        Cm = rng.normal(size=(nb, n, n))
        X = twin_run(twin, "lyap", Ast, Cm, n, nb * n * n).reshape(nb, n, n)
        res = Ast @ X + X @ Ast.transpose(0, 2, 1) + Cm
        assert np.max(np.abs(res)) < 1e-10
        Xr = np.array([sl.solve_continuous_lyapunov(a, -c) for a, c in zip(Ast[:50], Cm[:50])])
        assert np.max(np.abs(X[:50] - Xr) / np.max(np.abs(Xr), axis=(1, 2), keepdims=True)) < 1e-8


def test_twin_special_matrices(twin):
    # the reference's own cases (test/runtests.jl:497-523): rotation generator, identity Lyapunov, diagonal
    A = np.array([[[0.0, 1.0], [-1.0, 0.0]]])
    wr, wi = twin_run(twin, "eigvals", A, nout=2, im=True)
    assert np.array_equal(wr, [0.0, 0.0]) and np.array_equal(wi, [-1.0, 1.0])
    I2 = np.eye(2)[None]
    X = twin_run(twin, "lyap", I2, I2, 2, 4).reshape(2, 2)
    assert np.array_equal(X, sl.solve_continuous_lyapunov(np.eye(2), -np.eye(2)))      # -I/2
    E = twin_run(twin, "exp", np.zeros((1, 3, 3)), nout=9).reshape(3, 3)
    assert np.array_equal(E, np.eye(3))
    # repeated and defective eigenvalues
    J = np.array([[[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [0.0, 0.0, 2.0]]])
    wr, wi = twin_run(twin, "eigvals", J, nout=3, im=True)
    assert np.allclose(wr, 2.0, atol=1e-5) and np.allclose(wi, 0.0, atol=1e-5)
    # singular matrix: determinant exactly zero
    Z = np.array([[[1.0, 2.0], [2.0, 4.0]]])
    assert twin_run(twin, "det", Z, nout=1)[0] == 0.0
