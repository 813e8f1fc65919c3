"""Generate the committed golden fixtures under tests/golden/ (run here, in the build container):
    python tests/golden/make_golden.py

1. transcendental_truth.npz — seeded Float64 inputs and the CORRECTLY ROUNDED Float64 value of each
   hot-path transcendental computed with mpmath at 60 digits (the arbiter of the ulp tier; the
   reference's SLEEFPirates/Base implementations are not vendored).
2. reference_known_answers.json — the known-answer values published by the reference itself
   (README.md:31-35, src/MonteCarloMeasurements.jl:18-19, docs/src/index.md:55-62,73-74,346-354) with
   the file:line each one was read from.  These pin the oracle (tests/test_oracle_golden.py).
"""
import json
import os

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 60
rng = np.random.default_rng(20261019)


def rounded(f, xs):
    return np.array([float(f(mp.mpf(float(x)))) for x in xs])


n = 4000
out = {}
x_small = rng.normal(0, 2, n)
x_pi = np.pi + 0.1 * rng.normal(size=n)            # the C2 input distribution
x_big = rng.uniform(-1e4, 1e4, n)
x_pos = np.exp(rng.uniform(-30, 30, n))
x_unit = rng.uniform(-1, 1, n)
out["sin_x"] = np.concatenate([x_small, x_pi, x_big]); out["sin_y"] = rounded(mp.sin, out["sin_x"])
out["cos_x"] = out["sin_x"]; out["cos_y"] = rounded(mp.cos, out["cos_x"])
out["tan_x"] = np.concatenate([x_small, x_pi]); out["tan_y"] = rounded(mp.tan, out["tan_x"])
out["exp_x"] = np.concatenate([x_small, -x_pi, rng.uniform(-700, 700, n)]); out["exp_y"] = rounded(mp.exp, out["exp_x"])
out["log_x"] = np.concatenate([x_pos, np.abs(x_pi), 1 + 1e-3 * rng.normal(size=n)]); out["log_y"] = rounded(mp.log, out["log_x"])
out["expm1_x"] = np.concatenate([x_small, 1e-5 * x_small]); out["expm1_y"] = rounded(mp.expm1, out["expm1_x"])
out["log1p_x"] = np.concatenate([np.abs(x_small), 1e-5 * x_small]); out["log1p_y"] = rounded(mp.log1p, out["log1p_x"])
out["asin_x"] = x_unit; out["asin_y"] = rounded(mp.asin, x_unit)
out["acos_x"] = x_unit; out["acos_y"] = rounded(mp.acos, x_unit)
out["atan_x"] = np.concatenate([x_small, x_big]); out["atan_y"] = rounded(mp.atan, out["atan_x"])
out["sinh_x"] = x_small * 3; out["sinh_y"] = rounded(mp.sinh, out["sinh_x"])
out["cosh_x"] = x_small * 3; out["cosh_y"] = rounded(mp.cosh, out["cosh_x"])
out["tanh_x"] = x_small; out["tanh_y"] = rounded(mp.tanh, out["tanh_x"])
out["sqrt_x"] = x_pos; out["sqrt_y"] = rounded(mp.sqrt, x_pos)
pw_x = np.concatenate([np.abs(x_pi), x_pos[: n // 2] ** 0.1])
pw_y = np.concatenate([np.full(n, 1.5), rng.uniform(-8, 8, n // 2)])
out["pow_x"] = pw_x; out["pow_e"] = pw_y
out["pow_y"] = np.array([float(mp.power(mp.mpf(float(a)), mp.mpf(float(b)))) for a, b in zip(pw_x, pw_y)])
# inverse normal CDF on the systematic grid of N=2000 (config C1) and on random / tail probabilities
p_grid = (2 * np.arange(2000) + 1) / 4000.0
p_rand = np.concatenate([rng.random(2000), 10.0 ** (-rng.uniform(0, 300, 1000)), 1 - 10.0 ** (-rng.uniform(0, 15, 500))])
p_all = np.concatenate([p_grid, p_rand]); p_all = p_all[(p_all > 0) & (p_all < 1)]


def ninv(p):
    if p < mp.mpf("1e-8"):
        # erfinv loses the tail in mpmath: solve Phi(x) = p directly
        x0 = -mp.sqrt(-2 * mp.log(p))
        return mp.findroot(lambda x: mp.log(mp.erfc(-x / mp.sqrt(2)) / 2) - mp.log(p), x0)  # log-space: f ~ 1e-249 otherwise
    return mp.sqrt(2) * mp.erfinv(2 * p - 1) if p > mp.mpf("0.25") else -mp.sqrt(2) * mp.erfinv(1 - 2 * p)


out["norminv_p"] = p_all; out["norminv_y"] = rounded(ninv, p_all)
np.savez_compressed(os.path.join(HERE, "transcendental_truth.npz"), **out)

known = {
    "pstd_pi_pm_0.1": {"value": 0.09999231528930486, "source": "README.md:31-32; src/MonteCarloMeasurements.jl:18-19",
                       "what": "pstd(pi ± 0.1), N=2000 systematic Normal(pi,0.1)"},
    "static100_pstd": {"value": 0.9986403042113866, "source": "docs/src/index.md:55-59",
                       "what": "pstd(StaticParticles(100)), systematic Normal(0,1), N=100"},
    "static100_pvar": {"value": 0.9972824571954108, "source": "docs/src/index.md:55-59",
                       "what": "pvar(StaticParticles(100))"},
    "prob_lt_1": {"value": 0.8415, "source": "docs/src/index.md:346-354",
                  "what": "@prob Particles() < 1, N=2000: count 1683/2000"},
    "normal_fit_2x_plus_10": {"mu": 10.000000000000002, "sigma": 1.9972806084227737, "source": "docs/src/index.md:73-74",
                              "what": "Normal(f(p)), f(x)=2x+10, p=StaticParticles(100): (pmean, pstd)"},
    "two_mp_0.1_pstd_printed": {"value": 0.0999, "source": "README.md:28-30", "what": "2 ∓ 0.1 prints 2.0 ± 0.0999 (N=100)"},
}
json.dump(known, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1)
print("golden fixtures written:", {k: v.shape for k, v in out.items() if k.endswith("_y")})
