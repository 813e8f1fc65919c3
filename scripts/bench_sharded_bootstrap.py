"""Sharded bootstrap at C5 size: d variables x N particles (global), one shared index vector, sources read from their
owners' HBM over NVLink.  Launch: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 \
    --master-port 29511 scripts/bench_sharded_bootstrap.py [d] [N]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import mcm_b200 as mcm
mcm.sharded.init_distributed(rank, world, lr)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
lo, hi = mcm.sharded.shard_range(N, rank, world)
cols = mcm.systematic_batch(N, d, permute=True, seed=1, i0=lo, n_local=hi - lo) if False else None
# each rank builds its slice of d keyed systematic columns (global grid, no communication)
cols = [mcm.systematic_sample(N, mcm.Normal(0.0, 1.0), permute=False, i0=lo, n_total=N, n_local=hi - lo, as_particles=True) * (1.0 + j)
        for j in range(d)]
ts = []
for it in range(4):
    mcm.sync()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    out = mcm.sharded.bootstrap(cols, N, seed=3)
    mcm.sync()
    if world > 1: dist.barrier()
    ts.append(time.perf_counter() - t)
    if it < 3: del out
tt = torch.tensor([min(ts[1:])], device="cuda")
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
# check two variables against the global definition
idx = mcm.rand_indices(N, N, 3).to_numpy()[lo:hi]
grid = mcm.systematic_sample(N, mcm.Normal(0.0, 1.0), permute=False)
ok = all(np.array_equal(out[j].to_numpy(), (grid * (1.0 + j))[idx]) for j in (0, d - 1))
if rank == 0:
    t = tt.item()
    print(f"sharded bootstrap d={d} N={N} on {world} GPU(s): {t*1e3:.2f} ms  ({16*d*N/t/1e9:.0f} GB/s algorithmic aggregate, "
          f"{(world-1)/world*8*d*N/world/t/1e9:.0f} GB/s remote reads per GPU)  ok={ok}", flush=True)
mcm.sharded.free_windows()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
