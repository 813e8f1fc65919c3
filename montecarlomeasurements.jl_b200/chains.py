"""Canonical fused chains of the hot path, defined ONCE here and used by the
benchmarks, the parity tests and the ahead-of-time chain registry
(scripts/gen_chain_registry.py -> csrc/chain_registry.def)."""
from __future__ import annotations

from . import particles as P
from . import tape as T


def readme_expr(a):
    """README quick start: sin(a)*a^2 + exp(-a)   (config C1)."""
    return P.sin(a) * a ** 2 + P.exp(-a)


def chain10(a):
    """The 10-primitive headline chain (config C2):
    t1=sin(a); t2=a^2; t3=t1*t2; t4=-a; t5=exp(t4); t6=t3+t5; t7=abs(t6); t8=sqrt(t7); t9=a^1.5; y=t8/t9."""
    return P.sqrt(abs(P.sin(a) * a ** 2 + P.exp(-a))) / a ** 1.5


def pendulum_rhs(th, om, gL, c):
    """examples/ode_sim.jl:20-25 plus BASELINE's damping term: θ' = ω, ω' = -(g/L) sin θ - c ω."""
    return om, -(gL * P.sin(th)) - c * om


def make_rk4_step(dt: float):
    def step(th, om, gL, c):
        k1t, k1o = pendulum_rhs(th, om, gL, c)
        k2t, k2o = pendulum_rhs(th + (dt / 2) * k1t, om + (dt / 2) * k1o, gL, c)
        k3t, k3o = pendulum_rhs(th + (dt / 2) * k2t, om + (dt / 2) * k2o, gL, c)
        k4t, k4o = pendulum_rhs(th + dt * k3t, om + dt * k3o, gL, c)
        th2 = th + (dt / 6) * (k1t + 2.0 * k2t + 2.0 * k3t + k4t)
        om2 = om + (dt / 6) * (k1o + 2.0 * k2o + 2.0 * k3o + k4o)
        return th2, om2
    return step


def pendulum_rk4_tape(dt: float = 0.01) -> T.Tape:
    """In-register RK4: inputs (θ0, ω0, g, L, c), outputs (θ, ω) after `reps` steps (config C3)."""
    return T.trace_loop(make_rk4_step(dt), nstate=2, nparam=3, prologue=lambda g, Lp, c: (g / Lp, c))


REGISTERED = {
    "ReadmeExpr": lambda: T.trace(readme_expr, 1),
    "Chain10": lambda: T.trace(chain10, 1),
    "PendulumRK4": lambda: pendulum_rk4_tape(0.01),
}
