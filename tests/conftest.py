import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; oracle/mcm_oracle.c + oracle/oracle.py)."""
    import oracle as _o
    _o.build()
    return _o


@pytest.fixture(scope="session")
def golden():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "transcendental_truth.npz")))
    g["known"] = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")))
    return g


@pytest.fixture(scope="session")
def mcm():
    """The product package (device path).  GPU tests only."""
    import mcm_b200 as m
    m.init(0)
    return m


def ulp_err(got, ref):
    """|got - ref| in units of ulp(ref), in ref's own precision."""
    ref = np.asarray(ref)
    sp = np.spacing(np.abs(ref)).astype(np.float64)
    g = np.asarray(got, dtype=np.float64); r = ref.astype(np.float64)
    same = (g == r) | (np.isnan(g) & np.isnan(r))
    d = np.abs(g - r) / np.where(sp > 0, sp, 1.0)
    d[same] = 0.0
    d[np.isnan(d)] = np.inf
    return d


def bits_equal(a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    iv = np.int64 if a.dtype == np.float64 else np.int32
    an, bn = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(an, bn) and np.array_equal(a.view(iv)[~an], b.view(iv)[~bn]))
