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


def fused_stats(fn, *ps_local, materialize: bool = True):
    """`fused(fn)(p1, p2, ...; stats=True)` on particle-sharded variables: every rank evaluates the chain on its own
    slice (no communication) and receives the statistics of the GLOBAL result — the ranks' partials are exchanged and
    merged inside the kernel's epilogue over peer memory (mcm_chain_sharded).  Returns (local result or None, Moments).
    Collective."""
    from . import tape as _tape
    f = fn if isinstance(fn, P.fused) else P.fused(fn)
    ps = list(ps_local)
    N = ps[0].N
    dt = L.F64 if any(p.dtype == L.F64 for p in ps) else L.F32
    ps = [p.astype(P._NP[dt]) for p in ps]
    t = f.tape(len(ps))
    outs = [P._empty(N, dt) for _ in range(t.nout)] if materialize else []
    st = _tape.run_tape(t, [p.ptr for p in ps], N, dt, [o.ptr for o in outs] if materialize else None,
                        stats_output=0, sharded=True)
    res = (outs[0] if t.nout == 1 else tuple(outs)) if materialize else None
    return res, P.Moments(*st)


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


# ---- bootstrap across shards: P2P loads from the owners' windows ---------------------------------
_windows: dict[int, tuple[int, int]] = {}     # bytes -> (window id, local device pointer); windows are reused


def _window(nbytes: int):
    """Collective: every rank must ask for the same size in the same order."""
    nbytes = int(nbytes)
    if nbytes not in _windows:
        wid, ptr = C.c_int(-1), C.c_void_p()
        L.check(L.lib().mcm_window_create(nbytes, C.byref(wid), C.byref(ptr)))
        _windows[nbytes] = (wid.value, int(ptr.value))
    return _windows[nbytes]


def free_windows():
    """Collective: release every cached peer window."""
    for wid, _ in _windows.values():
        L.check(L.lib().mcm_window_free(wid))
    _windows.clear()


def bootstrap(p_local, N_global: int, n: int | None = None, *, seed: int = 0, force_window: bool = False):
    """bootstrap(p, n = N) (resampling.jl:76-88) on particle-sharded variables.

    p_local: this rank's slice (shard_range(N_global, rank, world)) of one Particles or of a list of Particles — ONE
    global index vector is applied to every variable (resampling.jl:80-84).  Output: this rank's slice
    (shard_range(n, rank, world)) of the resampled variables.  The indices are the same Philox stream as the
    single-process `bootstrap(p, n, seed=seed)`, so the sharded result is that result, sliced.  Source elements are read
    straight from their owners' HBM over NVLink (variables interleaved four to a 32-byte sector in the owners' windows);
    nothing else moves.  One rank: the ordinary shared-index gather (force_window=True exercises the window path)."""
    single = isinstance(p_local, P.Particles)
    ps = [p_local] if single else list(p_local)
    rank_, world = _comm["rank"], _comm["world"]
    N_global = int(N_global)
    n = N_global if n is None else int(n)
    lo, hi = shard_range(N_global, rank_, world)
    dt = ps[0].dtype
    if any(q.N != hi - lo or q.dtype != dt for q in ps):
        raise L.ArgumentError(L.MCM_ERR_INVALID, "sharded.bootstrap: every variable must hold this rank's slice of N_global particles")
    from . import resampling as R
    if world == 1 and not force_window:
        out = R.bootstrap(ps, n, seed=seed)
        return out[0] if single else out
    olo, ohi = shard_range(n, rank_, world)
    idx = R.rand_indices(ohi - olo, N_global, seed, offset=olo)
    outs = _gather_global(ps, idx, N_global, ohi - olo)
    return outs[0] if single else outs


def _gather_global(ps, idx, N_global: int, n_out_local: int):
    """outs[v][j] = variable v of GLOBAL particle idx[j]; ps = this rank's slices.  Collective."""
    rank_, world = _comm["rank"], _comm["world"]
    lo, hi = shard_range(N_global, rank_, world)
    dt = ps[0].dtype
    per = shard_range(N_global, 0, world)[1]                    # elements per rank (the last shard may be shorter)
    lib = L.lib()
    es = P._NP[dt].itemsize
    chunk = max(4, min(-(-len(ps) // 4) * 4, ((8 << 30) // max(per * es, 1)) // 4 * 4))    # variables per fill (<= 8 GB)
    wid, _ = _window(lib.mcm_window_bytes(chunk, per, dt))
    outs = P._empty_many(len(ps), n_out_local, dt)
    for v0 in range(0, len(ps), chunk):
        grp = ps[v0:v0 + chunk]
        sptr = (C.c_void_p * len(grp))(*[C.c_void_p(q.ptr) for q in grp])
        L.check(lib.mcm_window_fill(dt, wid, sptr, len(grp), hi - lo, per))
        L.check(lib.mcm_comm_fence())                                   # every owner has filled its window
        optr = (C.c_void_p * len(grp))(*[C.c_void_p(o.ptr) for o in outs[v0:v0 + chunk]])
        L.check(lib.mcm_gather_peer(dt, wid, len(grp), per, N_global, C.c_void_p(idx.ptr), n_out_local, optr))
        L.check(lib.mcm_comm_fence())                                   # every reader is done before the next fill
    return outs


def resample_indices(w_local, N_global: int, u: float = 0.5, n: int | None = None) -> P.DeviceIndex:
    """Systematic resampling indices for particle-sharded unnormalised weights: this rank's slice
    (shard_range(n, rank, world)) of the n GLOBAL source indices the single-process `resample_indices(weights, u)`
    returns.  Each rank scans its own weights; the run of outputs of every source is written straight into the index
    window of the rank that owns those output positions (P2P stores).  Collective."""
    rank_, world = _comm["rank"], _comm["world"]
    N_global = int(N_global)
    n = N_global if n is None else int(n)
    w = w_local if isinstance(w_local, P.Particles) else P.Particles(np.asarray(w_local, dtype=np.float64))
    w = w.astype(np.float64)
    lo, hi = shard_range(N_global, rank_, world)
    if w.N != hi - lo:
        raise L.ArgumentError(L.MCM_ERR_INVALID, "sharded.resample_indices: weights must be this rank's slice of N_global")
    per_out = max(shard_range(n, 0, world)[1], 1)
    olo, ohi = shard_range(n, rank_, world)
    wid, wptr = _window(per_out * 8 + (1 << 20))       # distinct size class from the data windows
    lib = L.lib()
    L.check(lib.mcm_comm_fence())                        # nobody is still reading an earlier result from the window
    L.check(lib.mcm_resample_systematic_sharded(C.c_void_p(w.ptr), w.N, N_global, lo, n, float(u), wid, per_out))
    L.check(lib.mcm_comm_fence())                        # every owner's runs have landed
    d = P.DeviceIndex(ohi - olo)
    if ohi > olo:
        L.check(lib.mcm_copy(C.c_void_p(d.ptr), C.c_void_p(wptr), (ohi - olo) * 8))
    return d


def resample(p_local, w_local, N_global: int, u: float = 0.5):
    """Weighted systematic resample of particle-sharded variables (the sharded form of `resample(p, weights, u)`):
    indices by `resample_indices`, values through the peer-window gather of `bootstrap`."""
    single = isinstance(p_local, P.Particles)
    ps = [p_local] if single else list(p_local)
    idx = resample_indices(w_local, N_global, u)
    outs = _gather_global(ps, idx, int(N_global), idx.n)
    return outs[0] if single else outs
