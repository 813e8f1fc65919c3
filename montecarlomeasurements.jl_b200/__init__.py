"""montecarlomeasurements.jl_b200 — B200-native particle arithmetic behind the
MonteCarloMeasurements.jl API (hot path only).

Import name: the directory name contains a dot, so import it through the shim
`mcm_b200` at the repository root (`import mcm_b200 as mcm`).

Everything device-side goes through libmcm_b200.so (include/mcm_b200.h); there
is no CPU implementation of any primitive in this package.
"""
from . import _lib
from ._lib import MCMError, ArgumentError, MethodError, init, is_initialised
from .particles import (
    Particles, Moments, DeviceIndex, ComparisonError, DomainError,
    DEFAULT_NUM_PARTICLES, DEFAULT_STATIC_NUM_PARTICLES,
    fill, fused, evalpoly, affine, lincomb,
    exp, exp2, exp10, expm1, log, log10, log2, log1p,
    sin, cos, tan, sind, cosd, tand, sinh, cosh, tanh,
    asin, acos, atan, asind, acosd, atand, asinh, acosh, atanh,
    zero, sign, sqrt, rad2deg, deg2rad, floor, ceil, trunc, inv, abs2, float_, round_,
    maximum, minimum, mod, mod1, rem, div, hypot, add_sum, clamp, muladd, minmax, sincos, ratdiv,
    moments, pmean, psum, pvar, pstd, pminimum, pmaximum, pextrema, pmiddle, piterate, meanstd, meanvar,
    pquantile, pmedian, psort, psortperm, pcov, pcor, to_matrix,
    pgemm, pgemv, mul_, pdot, paxpy_,
    unsafe_comparisons, set_comparison_function, approx, lesssim, gtrsim,
    check_domain, check_bounds, IntParticles, round_int, rem2pi, mod2pi, isinteger, iszero, isfinite, not_, fld, cld,
    RoundNearest, RoundToZero, RoundUp, RoundDown, RoundNearestTiesAway, RoundNearestTiesUp, RoundFromZero,
    nparticles, particleeltype, flat_particles, is_particle_array,
)
from .sampling import (Exponential, LogNormal, Cauchy, Laplace, Logistic, Rayleigh, Weibull, Pareto, Gumbel, TriangularDist,
                       Frechet, Kumaraswamy, Gamma, Erlang, Chisq, InverseGamma, Beta, TDist, FDist, GenericDistribution,
                       UnivariateDistribution, boxtimes, boxplus, outer_product, otimes, rand_particles_dist, quantile_dist)
from .sampling import (Normal, Uniform, MvNormal, systematic_sample, pm, mp, dotdot, transform_moments,
                       mvnormal_particles, systematic_batch, index_bijection, cholesky_lower, solve_right_lower, sigmapoints)
from .resampling import bootstrap, resample, resample_indices, rand_indices
from .optimize import optimize, perturb, wasserstein, ess, rand_particles
from .decision import union, intersect, with_nominal, nominal
from . import linalg, complexp
from .complexp import ComplexParticles, complex_particles
from .bymap import bymap, bypmap, prob, register_primitive, register_primitive_multi, register_primitive_single
from .deconstruct import has_particles, mean_object, replace_particles
from . import sharded, tape

StaticParticles = Particles  # storage is identical on device; N defaults differ (types.jl:26-30)


def sync():
    """Block until all queued device work is done."""
    _lib.check(_lib.lib().mcm_sync())
