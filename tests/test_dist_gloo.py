"""N>1 host-side logic on CPU: two gloo ranks shard the particle index, compute local partials with the
oracle, exchange them, and must reproduce the single-process oracle result with the same merge rules the
device path uses (six-moment Chan merge in rank order; centred Gram partials on the all-reduced mean)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, M, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mcm_b200 as mcm
    import oracle as orc
    lo, hi = mcm.sharded.shard_range(N, rank, world)
    # every rank generates ITS slice of the global systematic grid (no communication)
    a_loc = orc.systematic_normal(N, np.pi, 0.1, i0=lo, n_total=N, n_local=hi - lo)
    y_loc = orc.chain10(a_loc)
    part = torch.from_numpy(orc.moments(y_loc))
    gathered = [torch.zeros(6, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, part)
    acc = gathered[0].numpy()
    for r in range(1, world):
        acc = mcm.sharded.merge_moments(acc, gathered[r].numpy())
    # pcov: all-reduced sums -> global means; centred partial Gram; all-reduce
    rng = np.random.default_rng(5)
    X = rng.normal(size=(M, N)); X[2] = X[0] - 3 * X[1]
    Xl = X[:, lo:hi]
    sums = torch.from_numpy(np.concatenate([Xl.sum(axis=1), [hi - lo]]))
    dist.all_reduce(sums)
    mu = sums[:M].numpy() / sums[M].item()
    G = torch.from_numpy((Xl - mu[:, None]) @ (Xl - mu[:, None]).T)
    dist.all_reduce(G)
    if rank == 0:
        ret["moments"] = acc
        ret["cov"] = G.numpy() / (N - 1)
        ret["spans"] = (lo, hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_statistics_match_single_process(orc):
    N, M, world = 20001, 4, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, M, ret)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    a = orc.systematic_normal(N, np.pi, 0.1)
    want = orc.moments(orc.chain10(a))
    got = np.asarray(ret["moments"])
    assert got[0] == N
    assert abs(got[1] - want[1]) <= 1e-12 * abs(want[1])          # pmean
    assert abs(got[2] - want[2]) <= 1e-12 * want[2]               # M2 -> pstd
    assert got[3] == want[3] and got[4] == want[4]
    rng = np.random.default_rng(5)
    X = rng.normal(size=(M, N)); X[2] = X[0] - 3 * X[1]
    assert np.allclose(ret["cov"], orc.cov(X), rtol=1e-12, atol=1e-13)
