"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): one process per GPU, particle index sharded, only
reduction partials cross NVLink (NCCL inside libmcm_b200.so).  Results must equal the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, M, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import mcm_b200 as mcm
    from mcm_b200 import chains
    mcm.sharded.init_distributed(rank, world, rank)
    # elementwise + fused chain are rank-local; statistics are merged across ranks
    a = mcm.sharded.systematic_normal_shard(N, np.pi, 0.1)
    y, st = mcm.fused(chains.chain10)(a, stats=True)
    g = mcm.sharded.allmerge(st)                               # host-resident partials through NCCL
    gm = mcm.sharded.pmoments(y)                               # in-kernel exchange over peer memory (mcm_moments_sharded)
    y2, gk = mcm.sharded.fused_stats(chains.chain10, a)       # fused chain + in-kernel exchange (mcm_chain_sharded)
    same_y = bool(np.array_equal(y.to_numpy(), y2.to_numpy()))
    # deferred collection: three launches queued, statistics collected by ticket afterwards
    from mcm_b200 import tape as _tp
    t10 = _tp.trace(chains.chain10, 1)
    tickets = [_tp.run_tape(t10, [a.ptr], a.N, a.dtype, [y2.ptr], stats_output=0, sharded=True, wait=False) for _ in range(3)]
    later = [list(_tp.stats_wait(k)) for k in tickets]
    async_ok = all(l == later[0] for l in later) and later[0][:5] == [gk.n, gk.mean, gk.m2, gk.min, gk.max]
    # an unbalanced exchange: rank 0 contributes an empty shard
    e = mcm.Particles(np.zeros(0)) if rank == 0 else mcm.Particles(np.full(5, float(rank)))
    ge = mcm.sharded.pmoments(e)
    # sharded pcov of M variables
    rng = np.random.default_rng(5)
    X = rng.normal(1.0, 2.0, size=(M, N)); X[2] = X[0] - 3 * X[1]
    lo, hi = mcm.sharded.shard_range(N, rank, world)
    v = [mcm.Particles(np.ascontiguousarray(X[j, lo:hi])) for j in range(M)]
    Cg = mcm.sharded.pcov(v)
    # sharded quantiles: exact order statistics of the GLOBAL particle set; variable 3 is globally sorted, so each
    # bracket's candidates all live on one rank; variable 4 has heavy ties
    X[3] = np.sort(X[3]); X[4] = np.round(X[4])
    vq = [mcm.Particles(np.ascontiguousarray(X[j, lo:hi])) for j in range(M)]
    Q = mcm.sharded.pquantile(vq, [0.025, 0.5, 0.975])
    Q5 = mcm.sharded.pquantile(vq[0], [0.0, 0.1, 0.5, 0.9, 1.0])
    qy = mcm.sharded.pquantile(y, 0.5)
    # sharded bootstrap: one global index vector, sources read from their owners over NVLink; equals the
    # single-process bootstrap (same Philox stream), sliced
    boot_ok = True
    for n_out, seed in ((N, 3), (N // 2 + 3, 4)):
        outs = mcm.sharded.bootstrap(v[:3], N, n_out, seed=seed)
        full_idx = mcm.rand_indices(n_out, N, seed).to_numpy()
        olo, ohi = mcm.sharded.shard_range(n_out, rank, world)
        for j in range(3):
            boot_ok &= bool(np.array_equal(outs[j].to_numpy(), X[j][full_idx[olo:ohi]]))
    one = mcm.sharded.bootstrap(vq[3], N, seed=9)
    full_idx = mcm.rand_indices(N, N, 9).to_numpy()
    boot_ok &= bool(np.array_equal(one.to_numpy(), X[3][full_idx[lo:hi]]))
    x32 = mcm.Particles(np.ascontiguousarray(X[0, lo:hi].astype(np.float32)))
    o32 = mcm.sharded.bootstrap(x32, N, seed=9)
    boot_ok &= bool(np.array_equal(o32.to_numpy(), X[0].astype(np.float32)[full_idx[lo:hi]]))
    # sharded systematic resampling: identical to the single-process indices (same fixed-point scale), sliced;
    # weights with zeros, a run longer than a warp job and one dominant weight (queued fill path)
    wrng = np.random.default_rng(11)
    wfull = wrng.random(N) ** 4
    wfull[wrng.integers(0, N, N // 10)] = 0.0
    wfull[N // 3] = 5e3
    wfull[N - 7] = 0.6 * wfull.sum()
    res_ok = True
    for uu in (0.5, 0.0, 0.999):
        want_idx = mcm.resample_indices(wfull, uu).to_numpy()
        got = mcm.sharded.resample_indices(np.ascontiguousarray(wfull[lo:hi]), N, uu)
        olo, ohi = mcm.sharded.shard_range(N, rank, world)
        res_ok &= bool(np.array_equal(got.to_numpy(), want_idx[olo:ohi]))
    rs = mcm.sharded.resample(v[:3], np.ascontiguousarray(wfull[lo:hi]), N, 0.999)
    for j in range(3):
        res_ok &= bool(np.array_equal(rs[j].to_numpy(), X[j][want_idx[olo:ohi]]))
    boot_ok &= res_ok
    mcm.check_domain()
    mcm.sharded.free_windows()
    ret[rank] = {"fused": [g.n, g.mean, g.m2, g.min, g.max], "pm": [gm.n, gm.mean, gm.m2, gm.min, gm.max],
                 "kernel": [gk.n, gk.mean, gk.m2, gk.min, gk.max], "same_y": same_y, "async_ok": async_ok,
                 "empty": [ge.n, ge.mean, ge.min, ge.max], "cov": Cg,
                 "local_n": a.N, "Q": Q, "Q5": Q5, "qy": qy, "boot_ok": boot_ok}
    dist.barrier()
    dist.destroy_process_group()


def _worlds():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    # every power of two the box offers (8 ranks included) and one odd world (quantile sample split, shard ranges)
    return [w for w in (2, 3, 4, 8) if w <= n]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_statistics_and_pcov_match_oracle(orc, world):
    import torch
    import torch.multiprocessing as mp
    if world not in _worlds():
        pytest.skip(f"needs >= {world} GPUs")
    N, M = 2_000_001, 6
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, M, ret)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(500)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    a = orc.systematic_normal(N, np.pi, 0.1)
    want = orc.moments(orc.chain10(a))
    assert sum(ret[r]["local_n"] for r in range(world)) == N
    for r in range(world):
        for key in ("fused", "pm", "kernel"):
            got = ret[r][key]
            assert got[0] == N
            assert abs(got[1] - want[1]) <= 1e-12 * abs(want[1]) and abs(got[2] - want[2]) <= 1e-11 * want[2], (r, key)
        assert ret[r]["fused"] == ret[0]["fused"]                  # identical on every rank
        assert ret[r]["kernel"] == ret[0]["kernel"] and ret[r]["pm"] == ret[0]["pm"]   # in-kernel merge: bit-identical
        assert ret[r]["same_y"] and ret[r]["async_ok"]
        ne = 5 * (world - 1)
        assert ret[r]["empty"] == [ne, sum(5.0 * k for k in range(1, world)) / ne, 1.0, float(world - 1)]
        assert ret[r]["boot_ok"], r
    rng = np.random.default_rng(5)
    X = rng.normal(1.0, 2.0, size=(M, N)); X[2] = X[0] - 3 * X[1]
    Cr = orc.cov(X)
    for r in range(world):
        assert np.max(np.abs(ret[r]["cov"] - Cr)) <= 1e-12 * np.max(np.abs(Cr))
        assert np.array_equal(ret[r]["cov"], ret[0]["cov"])
    X[3] = np.sort(X[3]); X[4] = np.round(X[4])
    for r in range(world):
        for j in range(M):
            assert np.array_equal(ret[r]["Q"][j], orc.quantile(X[j], [0.025, 0.5, 0.975])), (r, j)   # bit-exact, every rank
        assert np.array_equal(ret[r]["Q5"], orc.quantile(X[0], [0.0, 0.1, 0.5, 0.9, 1.0]))
        assert ret[r]["qy"] == orc.quantile(orc.chain10(a), 0.5)[0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multi_gpu_parity.log"), "a") as fh:
        fh.write("world=%d ranks: moments / fused-chain statistics (NCCL path and in-kernel peer exchange) within 1e-12 of the "
                 "single-process oracle and bit-identical across ranks; pcov 1e-12; quantiles, bootstrap, resample bit-exact: PASSED\n" % world)
