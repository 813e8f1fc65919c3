"""Measure, on a B200, how far every ulp-tier primitive is from (a) the CPU oracle (glibc) and (b) the correctly
rounded truth (mpmath, Float64 only).  Output: gpurun_out/ulps.txt.  Used to set the limits in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpmath as mp
import mcm_b200 as mcm
import oracle as orc
from conftest import ulp_err

mp.mp.dps = 40
mcm.init(0)
N = 200000
rng = np.random.default_rng(1)
x = rng.normal(0.3, 1.0, N)
DOM = {
    "EXP": lambda x: x, "EXP2": lambda x: x, "EXP10": lambda x: x, "EXPM1": lambda x: x,
    "LOG": lambda x: np.abs(x) + 0.01, "LOG10": lambda x: np.abs(x) + 0.01, "LOG2": lambda x: np.abs(x) + 0.01, "LOG1P": np.abs,
    "SIN": lambda x: x * 3, "COS": lambda x: x * 3, "TAN": lambda x: x, "SINH": lambda x: x * 2, "COSH": lambda x: x * 2, "TANH": lambda x: x,
    "ASIN": lambda x: np.clip(x, -0.99, 0.99), "ACOS": lambda x: np.clip(x, -0.99, 0.99), "ATAN": lambda x: x * 5,
    "ASIND": lambda x: np.clip(x, -0.99, 0.99), "ACOSD": lambda x: np.clip(x, -0.99, 0.99), "ATAND": lambda x: x * 5,
    "ASINH": lambda x: x * 5, "ACOSH": lambda x: np.abs(x) + 1.01, "ATANH": lambda x: np.clip(x, -0.99, 0.99),
    "SIND": lambda x: x * 100, "COSD": lambda x: x * 100, "TAND": lambda x: x * 40,
}
MP = {"EXP": mp.exp, "EXP2": lambda v: mp.power(2, v), "EXP10": lambda v: mp.power(10, v), "EXPM1": mp.expm1, "LOG": mp.log,
      "LOG10": mp.log10, "LOG2": lambda v: mp.log(v, 2), "LOG1P": mp.log1p, "SIN": mp.sin, "COS": mp.cos, "TAN": mp.tan,
      "SINH": mp.sinh, "COSH": mp.cosh, "TANH": mp.tanh, "ASIN": mp.asin, "ACOS": mp.acos, "ATAN": mp.atan,
      "ASIND": lambda v: mp.degrees(mp.asin(v)), "ACOSD": lambda v: mp.degrees(mp.acos(v)), "ATAND": lambda v: mp.degrees(mp.atan(v)),
      "ASINH": mp.asinh, "ACOSH": mp.acosh, "ATANH": mp.atanh,
      "SIND": lambda v: mp.sin(mp.radians(v)), "COSD": lambda v: mp.cos(mp.radians(v)), "TAND": lambda v: mp.tan(mp.radians(v))}
lines = []
for name, dom in DOM.items():
    row = [name]
    for dt in (np.float64, np.float32):
        xi = np.ascontiguousarray(dom(x).astype(dt))
        got = mcm.particles._unary(name, mcm.Particles(xi)).to_numpy()
        row.append("%s vs oracle %.2f" % (dt.__name__, ulp_err(got, orc.unary(name, xi)).max()))
        k = 3000
        truth = np.array([float(MP[name](mp.mpf(float(v)))) for v in xi[:k]]).astype(dt) if dt == np.float64 else \
            np.array([float(MP[name](mp.mpf(float(v)))) for v in xi[:k]], dtype=np.float64).astype(np.float32)
        row.append("vs truth %.2f" % ulp_err(got[:k], truth).max())
        row.append("oracle vs truth %.2f" % ulp_err(orc.unary(name, xi)[:k], truth).max())
    lines.append(" | ".join(row))
y = rng.normal(-0.2, 2.0, N)
for name, f, mpf in (("ATAN2", mcm.atan, mp.atan2), ("HYPOT", mcm.hypot, mp.hypot), ("POW", None, mp.power)):
    row = [name]
    for dt in (np.float64, np.float32):
        a = x.astype(dt); b = y.astype(dt)
        if name == "POW":
            a = (np.abs(a) + dt(0.01)).astype(dt)
            got = (mcm.Particles(a) ** mcm.Particles(b)).to_numpy()
        else:
            got = f(mcm.Particles(a), mcm.Particles(b)).to_numpy()
        ref = orc.binary(name, a, b)
        row.append("%s vs oracle %.2f" % (dt.__name__, ulp_err(got, ref).max()))
        k = 3000
        truth = np.array([float(mpf(mp.mpf(float(u)), mp.mpf(float(v)))) for u, v in zip(a[:k], b[:k])], dtype=np.float64).astype(dt)
        row.append("vs truth %.2f" % ulp_err(got[:k], truth).max())
        row.append("oracle vs truth %.2f" % ulp_err(ref[:k], truth).max())
    lines.append(" | ".join(row))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "ulps.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
