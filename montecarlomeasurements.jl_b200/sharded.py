"""Multi-GPU execution: one process per GPU, the particle index sharded across ranks.

Rank r of g holds the contiguous slice [lo, hi) of EVERY uncertain variable
(index alignment across variables is what makes a vector of Particles a joint
distribution, src/types.jl:117-118).  Elementwise primitives and fused chains run
locally with no communication; only reduction partials cross NVLink (NCCL inside
libmcm_b200.so): six moments per rank for pmean/pstd/pextrema, M sums and one
M x M Gram block for pcov.  The reference has no distributed path (`bypmap` is a
plain `map`, src/bymap.jl:90-108).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from . import particles as P
from . import sampling as S


def shard_range(N: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one, 2-element aligned
    boundaries so every shard keeps 16-byte vector alignment for Float64."""
    N, rank, world = int(N), int(rank), int(world)
    if not (0 <= rank < world):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "rank out of range")
    per = -(-N // world)
    per += per % 2
    lo = min(N, rank * per)
    hi = min(N, lo + per)
    return lo, hi


def merge_moments(a, b):
    """Chan/Golub/LeVeque merge of two {n, mean, M2, min, max, sum} partials (host mirror of
    mcm_moments_merge; used to check the device merge and by the gloo tests)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if b[0] == 0: return a.copy()
    if a[0] == 0: return b.copy()
    n = a[0] + b[0]
    d = b[1] - a[1]
    fb = b[0] / n
    mn = a[3] if np.isnan(a[3]) else (b[3] if np.isnan(b[3]) else min(a[3], b[3]))
    mx = a[4] if np.isnan(a[4]) else (b[4] if np.isnan(b[4]) else max(a[4], b[4]))
    return np.array([n, a[1] + (0.0 + d * fb), (a[2] + b[2]) + (d * d) * (a[0] * fb), mn, mx, a[5] + b[5]])


_comm = {"rank": 0, "world": 1, "ready": False}


def init_distributed(rank: int | None = None, world: int | None = None, local_rank: int | None = None):
    """Bind to GPU LOCAL_RANK and create the library's NCCL communicator.  The 128-byte NCCL unique id is
    created on rank 0 and distributed through torch.distributed (which must be initialised; any backend)."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank))) if local_rank is None else local_rank
    lib = L.init(local_rank)
    _comm.update(rank=rank, world=world)
    if world > 1 and not _comm["ready"]:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise L.MCMError(L.MCM_ERR_NCCL, "init_distributed: torch.distributed must be initialised first")
        ident = C.create_string_buffer(128)
        if rank == 0:
            L.check(lib.mcm_comm_unique_id(ident))
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0)
        ident = C.create_string_buffer(box[0], 128)
        L.check(lib.mcm_comm_init(ident, world, rank))
        _comm["ready"] = True
    return rank, world


def rank() -> int: return _comm["rank"]
def world_size() -> int: return _comm["world"]


def systematic_normal_shard(N_global: int, mu: float, sigma: float, dtype=np.float64):
    """This rank's slice of systematic_sample(N_global, Normal(mu,sigma); permute=false) — generated
    locally from the global grid indices, no communication."""
    lo, hi = shard_range(N_global, _comm["rank"], _comm["world"])
    d = S.Normal(dtype(mu), dtype(sigma))
    return S.systematic_sample(N_global, d, permute=False, i0=lo, n_total=N_global, n_local=hi - lo,
                               as_particles=True)


def pmoments(p_local: P.Particles) -> P.Moments:
    """Global {n, mean, M2, min, max, sum} of a sharded Particles (identical on all ranks)."""
    m = np.zeros(6)
    L.check(L.lib().mcm_moments_sharded(p_local.dtype, C.c_void_p(p_local.ptr), p_local.N, m.ctypes.data_as(L._dp)))
    return P.Moments(*m)


def allmerge(local: P.Moments) -> P.Moments:
    """Merge host-resident per-rank moments (the fused statistics of a chain) across ranks."""
    m = np.array([local.n, local.mean, local.m2, local.min, local.max, local.sum])
    L.check(L.lib().mcm_moments_allmerge(m.ctypes.data_as(L._dp)))
    return P.Moments(*m)


def pmean(p_local): return pmoments(p_local).mean
def pstd(p_local, corrected=True): return pmoments(p_local).std(corrected)
def pvar(p_local, corrected=True): return pmoments(p_local).var(corrected)
def pextrema(p_local):
    m = pmoments(p_local)
    return m.min, m.max


def pcov(v_local, corrected: bool = True) -> np.ndarray:
    """Global pcov of M sharded variables: all-reduced means, local centred Gram, all-reduced M x M block."""
    v = P._check_mv(v_local)
    M = len(v)
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
    dC = P._alloc(M * M * 8)
    try:
        L.check(L.lib().mcm_gram_sharded(v[0].dtype, ptrs, M, v[0].N, 1 if corrected else 0, C.c_void_p(dC)))
        out = np.empty((M, M))
        L.check(L.lib().mcm_download(out.ctypes.data_as(C.c_void_p), C.c_void_p(dC), out.nbytes))
    finally:
        L.lib().mcm_free(C.c_void_p(dC))
    return out


def pquantile(p_local, q):
    """Global pquantile of a sharded Particles (or a list of M sharded variables): every rank passes its shard and
    gets the quantiles of the whole particle set (mcm_quantile_sharded: sampled pivots, local count/collect, radix
    select on all-reduced histograms; only <= 2048 finalists per rank group cross NVLink)."""
    import numbers
    scalar = isinstance(q, numbers.Real)
    qs = np.ascontiguousarray(np.atleast_1d(np.asarray(q, dtype=np.float64)))
    single = isinstance(p_local, P.Particles)
    v = [p_local] if single else P._check_mv(p_local)
    M = len(v)
    ptrs = (C.c_void_p * M)(*[C.c_void_p(x.ptr) for x in v])
    out = np.zeros((M, len(qs)))
    L.check(L.lib().mcm_quantile_sharded(v[0].dtype, ptrs, M, v[0].N, qs.ctypes.data_as(L._dp), len(qs),
                                         out.ctypes.data_as(L._dp)))
    if single:
        return float(out[0, 0]) if scalar else out[0]
    return out[:, 0].copy() if scalar else out


def pmedian(p_local):
    """Global median = middle of the one / two central order statistics of the whole particle set.  (Computed as the
    type-7 quantile at 0.5, which interpolates the same two order statistics: a + (b - a)/2.)"""
    return pquantile(p_local, 0.5)
