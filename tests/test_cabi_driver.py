"""The C++ caller of the C ABI (tests/cabi/cabi_driver.cpp; SURVEY.md §8b "a C++ test driver"): built against
include/mcm_b200.h + libmcm_b200.so with plain g++, no Python or torch in the process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "montecarlomeasurements.jl_b200")
SRC = os.path.join(ROOT, "tests", "cabi", "cabi_driver.cpp")
BUILD = os.path.join(ROOT, "tests", "cabi", "_build")
EXE = os.path.join(BUILD, "cabi_driver")


def build():
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["g++", "-O1", "-std=c++17", "-o", EXE, SRC, "-L" + PKG, "-lmcm_b200", "-Wl,-rpath," + PKG]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return EXE


def test_driver_builds_and_refuses_to_run_without_a_gpu():
    exe = build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)          # loud failure, no CPU fallback
    assert "mcm_init failed" in p.stdout


@pytest.mark.gpu
def test_driver_runs_the_quick_start_through_the_c_abi():
    exe = build()
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout, p.stderr)
    assert "all checks passed" in p.stdout
