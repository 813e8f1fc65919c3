"""optimize, wasserstein, ess — the consumers of the device reductions that SURVEY.md §8(f) ranks next
(src/optimize.jl:3-30, src/distances.jl:7-13, src/sampling.jl:26-38).  Every step runs on device through
the C ABI: weights exp.(τ y) (elementwise), weighted resampling (integer prefix scan), shared-index gather,
pcov (Gram), and the re-perturbation μ + chol(1.1 Cp + 1e-12 I) z with iid z (mcm_rand_real + mcm_pgemm).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import particles as P
from . import resampling as R
from . import sampling as S


def rand_particles(N: int, d=None, *, seed: int = 0, offset: int = 0, dtype=np.float64) -> P.Particles:
    """Particles(rng, N, d; systematic=false) (types.jl:67-82): N iid draws from Normal / Uniform on device
    (counter-based Philox stream keyed by (seed, offset); Julia's rng stream is not reproducible, H5)."""
    d = S.Normal() if d is None else d
    code = P._dtype_code(dtype)
    out = P._empty(int(N), code)
    if isinstance(d, S.Normal):
        kind, a, b = 0, float(d.mu), float(d.sigma)
    elif isinstance(d, S.Uniform):
        kind, a, b = 1, float(d.a), float(d.b)
    else:
        raise L.MethodError(L.MCM_ERR_UNSUPPORTED, "rand_particles supports Normal and Uniform")
    L.check(L.lib().mcm_rand_real(code, kind, C.c_void_p(out.ptr), int(N), a, b, int(seed), int(offset)))
    return out


def perturb(p, Cp, seed: int = 0):
    """perturb(rng, p, Cp) (optimize.jl:3-6): Particles(rng, N, MvNormal(pmean(p), 1.1Cp + 1e-12 I)) = μ + L z with
    iid standard normal z (types.jl:92-100), L = cholesky(1.1Cp + 1e-12 I).L (device potrf), applied by mcm_pgemm."""
    p = P._check_mv(p)
    d, N = len(p), p[0].N
    mu = np.array([P.pmean(x) for x in p])
    Lc = S.cholesky_lower(1.1 * np.asarray(Cp) + 1e-12 * np.eye(d))
    z = [rand_particles(N, seed=seed, offset=j * N, dtype=p[0].np_dtype) for j in range(d)]
    return P.pgemm(Lc, z, bias=mu)


def optimize(f, p, tau: float = 1.0, *, tau_i: float = 1.005, iters: int = 10000, tol: float = 1e-8, seed: int = 0,
             resampling: str = "systematic", info=None, trace=None):
    """res = optimize([rng,] f, p, τ=1; τi=1.005, iters=10000, tol=1e-8) (optimize.jl:14-30): simulated-annealing
    style minimisation of f over a particle cloud.  Per iteration: y = -f(p); we = exp.(τ y); j = weighted sample of
    N ancestors (ordered); every variable gathered with the SAME j; Cp = pcov(p); stop when tr(Cp) < tol; else
    re-perturb around the mean and raise τ.  `resampling`: "systematic" (exact integer prefix scan, one uniform
    offset per iteration) stands in for StatsBase.sample(…, ProbabilityWeights, ordered=true).
    Returns the final list of Particles (a new list; the input is not modified, like the reference's deepcopy).
    `trace` (a list, tests only): every iteration appends the host copies of its intermediates (weights, the uniform
    offset, ancestors, gathered cloud, Cp, the perturbation seed, the new cloud, τ) so that each step can be replayed
    against the oracle with the same Philox stream."""
    p = [x for x in P._check_mv(p)]
    N = p[0].N
    rng = np.random.default_rng(seed)
    tau = float(tau)
    for it in range(int(iters)):
        y = -f(p)
        if not isinstance(y, P.Particles):
            raise L.ArgumentError(L.MCM_ERR_INVALID, "optimize: f must return a Particles")
        we = P.exp(tau * y)
        u = float(rng.random())
        j = R.resample_indices(we, u=u)
        p = R.bootstrap(p, N, inds=j)
        Cp = P.pcov(p)
        if trace is not None:
            trace.append({"tau": tau, "y": y.to_numpy(), "we": we.to_numpy(), "u": u, "j": j.to_numpy(),
                          "gathered": [x.to_numpy() for x in p], "Cp": np.array(Cp)})
        if np.trace(Cp) < tol:
            if info is not None:
                info["iterations"] = it + 1
                info["converged"] = True
            return p
        pseed = int(rng.integers(0, 2**62))
        p = perturb(p, Cp, seed=pseed)
        if trace is not None:
            trace[-1].update(seed=pseed, perturbed=[x.to_numpy() for x in p])
        tau *= tau_i
    if info is not None:
        info["iterations"] = int(iters)
        info["converged"] = False
    return p


def wasserstein(p1: P.Particles, p2: P.Particles, p) -> float:
    """wasserstein(p1, p2, p) (distances.jl:7-13): mean_i |sort(p1)_i - sort(p2)_i|^p (W_p to the p-th power)."""
    if p1.N != p2.N:
        raise L.MethodError(L.MCM_ERR_INVALID, "wasserstein: both arguments need the same number of particles")
    d = abs(P.psort(p1) - P.psort(p2))
    return float(P.pmean(d if p == 1 else d ** p))


def ess(p: P.Particles) -> float:
    """ess(p) (sampling.jl:26-38): effective sample size from the autocorrelation at lags 1..min(250, N÷2)
    (Geyer's initial positive sequence, as MCMCDiagnostics.jl)."""
    N = p.N
    nl = min(250, N // 2)
    if nl < 1:
        raise L.ArgumentError(L.MCM_ERR_INVALID, "ess: need at least 2 particles")
    ac = np.zeros(nl)
    L.check(L.lib().mcm_autocor(p.dtype, C.c_void_p(p.ptr), N, nl, ac.ctypes.data_as(L._dp)))
    n = len(ac)
    tau_inv = 1 + 2 * ac[0]
    K = 2
    while K < n - 2:                      # 1-based K in the reference: Δ = ac[K] + ac[K+1]
        delta = ac[K - 1] + ac[K]
        if delta < 0:
            break
        tau_inv += 2 * delta
        K += 2
    return float(min(1 / tau_inv, 1.0) * N)
