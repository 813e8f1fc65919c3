"""Build libmcm_b200.so (the C-ABI CUDA library) in-tree for sm_100a.

Usage: python montecarlomeasurements.jl_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to csrc/build/, the shared
library next to this file (git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libmcm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false: `+ - * /` must round exactly like the reference's separate
# broadcasts (BASELINE.json north_star); fused multiply-adds are written
# explicitly (fma()) where the algorithm wants them.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "-diag-suppress", "177,550",
]
# experiments: e.g. MCM_NVCC_EXTRA="-DMCM_CHAIN_W=2 -DMCM_CHAIN_MIN_CTAS=2"
NVCC_FLAGS += os.environ.get("MCM_NVCC_EXTRA", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_hash(src: str) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in sorted(os.listdir(CSRC)):
        # every non-.cu file is a potential include (.cuh, .h, .inc coefficient tables, the generated .def registry)
        if f == src or (os.path.isfile(os.path.join(CSRC, f)) and not f.endswith(".cu")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "mcm_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".hash"
    want = _deps_hash(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    if verbose:
        sys.stderr.write(p.stderr)
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-ldl", "-lpthread"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
