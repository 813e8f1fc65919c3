#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json on 1..8 B200.

Workload: one uncertain scalar, Float64 particles sharded over the GPUs on the particle
index, a = pi + 0.1 * Phi^-1((2i+1)/(2N)) (the systematic sample of `pi ± 0.1`); a step is ONE pass of the hot
path: the canonical 10-primitive chain (sin, ^2, *, neg, exp, +, abs, sqrt, ^1.5, /) evaluated as one fused
launch that materialises y and reduces pmean(y)/pstd(y) in the same pass; for N > 1 ranks the kernel's last CTA
exchanges the rank's partial moments with every peer over NVLink (P2P stores into peer windows) and merges them,
and the 48-byte GLOBAL result is written straight into mapped pinned host memory.  The host collects the result of
step k while kernel k+1 runs (all K results are collected inside the timed region).
metric = Gparticle-ops/s = 10 * particles / time, whole job.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (weak scaling: 1e8 particles per GPU)
    python bench.py --gpus N --scaling strong                # north_star's literal split: 1e8 particles in total, N/g per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU execution model (stand-in)

One JSON line on stdout (rank 0).  The weak-scaling line also carries the strong-scaling measurement
(`strong_scaling`), C3 (`c3_pendulum`), C4 (`pcov`), C5 (`c5_pipeline`) and the HBM-bound family.
See DESIGN.md §6 for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 100_000_000
OPS_PER_PARTICLE = 10
ALGO_BYTES_PER_PARTICLE = 16          # read a (8 B) + write y (8 B); statistics are fused
CPU_SAMPLE_N = 100_000_000            # the CPU arm runs the SAME N per step (one pass is ~0.25 s on 16 threads)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Sample SM clocks and throttle reasons DURING the timed region (NVML, 2 ms period)."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = get(self.h)
                for b, nm in names.items():
                    if bits & b:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(np.min(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "period_ms": 1e3 * self.period}


# ----------------------------------------------------------------------------------------------------
def cpu_reference_arm(steps: int, warmup: int, threads: int | None = None, n: int = CPU_SAMPLE_N):
    """The reference's CPU execution model, timed on this box's host cores (oracle/ref_cpu.cpp: one SIMD
    pass per primitive with SLEEF kernels, temporaries materialised, mean 1 pass + std 2 passes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    # compile for THIS host (the .so shipped from the build container may target another CPU)
    orc.build(force=True)
    r = orc.refcpu()
    threads = threads or r.ref_max_threads()
    try:
        a = orc.systematic_normal(n, np.pi, 0.1)
        orc.ref_chain10(a[:1024], threads=1)
        _ = orc.ref_chain10(a, threads=threads)          # allocates and touches the work vectors (4 x 8 n bytes)
    except MemoryError:
        n = n // 10
        a = orc.systematic_normal(n, np.pi, 0.1)
    times = []
    res = None
    for i in range(warmup + steps):
        y, m, s, dt = orc.ref_chain10(a, threads=threads)
        res = (m, s)
        if i >= warmup:
            times.append(dt)
    n_eff = (n // r.ref_vector_width()) * r.ref_vector_width()
    total = sum(times)
    value = OPS_PER_PARTICLE * n_eff * len(times) / total / 1e9
    return {"value": value, "ms_per_step": 1e3 * total / len(times), "threads": threads, "n": n_eff,
            "mean": res[0], "std": res[1], "simd_width": r.ref_vector_width()}


def print_reference_line(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    steps = min(args.steps, 40)       # bounded: one step is a full N = 1e8 pass (~0.25 s on 16 threads)
    r = cpu_reference_arm(steps, min(args.warmup, 3))
    sample = (f"chain10 + pmean/pstd on N={r['n']} Float64 particles per step (the per-GPU workload of the CUDA arm), "
              f"{steps} timed steps, SLEEF 3.8 AVX{'-512' if r['simd_width'] == 8 else '2'} loop-per-primitive port of the "
              f"reference's execution model, persistent pool of {r['threads']} host threads; the Julia reference itself cannot "
              f"run here (no julia binary)")
    line = {
        "impl": "reference", "metric": "Gparticle-ops/s", "value": r["value"], "unit": "Gparticle-ops/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: single uncertain scalar, N=1e8 Float64 particles, 10-primitive chain + pmean/pstd "
                               "(CPU arm: the same N per step, one host)", "particles_per_step": r["n"]},
        "cpu_baseline": {"value": r["value"], "unit": "Gparticle-ops/s", "cores": r["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["value"], "unit": "Gparticle-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "check": {"pmean": r["mean"], "pstd": r["std"]},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
class C2Runner:
    """chain10 + fused statistics on this rank's slice [lo, lo + n) of the n_total-point global grid."""

    def __init__(self, mcm, n_local, n_total, lo, world, stream):
        from mcm_b200 import chains, tape as tp
        self.mcm, self.L, self.lib = mcm, mcm._lib, mcm._lib.lib()
        self.n, self.world, self.stream = n_local, world, stream
        L = self.L
        self.t10 = tp.trace(chains.chain10, 1)
        self.a = mcm.systematic_sample(n_total, mcm.Normal(np.pi, 0.1), permute=False, i0=lo, n_total=n_total,
                                       n_local=n_local, as_particles=True)
        self.y = mcm.particles._empty(n_local, L.F64)
        self.ins = (C.c_void_p * 1)(self.a.ptr)
        self.outs = (C.c_void_p * 1)(self.y.ptr)
        self.consts = np.asarray(self.t10.consts)
        self.tape_c = self.t10.as_ctypes()
        self.lb, self.le = self.t10.bounds
        self.fn = self.lib.mcm_chain_sharded if world > 1 else self.lib.mcm_chain
        self.buf = np.zeros(6)

    def launch(self, stats=True):
        """asynchronous: returns the ticket of the launch"""
        t = self.t10
        self.L.check(self.fn(self.L.F64, self.tape_c, len(t.ins), self.lb, self.le, 0, self.consts.ctypes.data_as(self.L._dp),
                             len(t.consts), self.ins, 1, self.outs, 1, self.n, t.out_slots[0] if stats else -1, None))
        return self.lib.mcm_chain_ticket() if stats else None

    def collect(self, ticket):
        self.L.check(self.lib.mcm_chain_stats_wait(ticket, self.buf.ctypes.data_as(self.L._dp)))
        return self.mcm.Moments(*self.buf)

    def steps(self, k):
        """k pipelined steps: the result of step j is collected while kernel j+1 runs; all k are collected."""
        prev, m = None, None
        for _ in range(k):
            t = self.launch()
            if prev is not None:
                m = self.collect(prev)
            prev = t
        return self.collect(prev)


def timed_c2(run: C2Runner, steps, warmup, barrier, torch, dist, world, local_rank, sample_clocks=True):
    stream = run.stream
    with torch.cuda.stream(stream):
        m = run.steps(warmup)
        barrier()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(local_rank); sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        m = run.steps(steps)
        ev1.record(stream)
        barrier()
        clocks = sampler.result() if sampler else None
        elapsed_ms = ev0.elapsed_time(ev1)
        # kernel-only duration for the roofline: the same launches queued back to back, nothing collected
        kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kreps = max(20, min(steps, 100))
        barrier()
        kev0.record(stream)
        for _ in range(kreps):
            run.launch()
        kev1.record(stream)
        torch.cuda.synchronize()
        kernel_ms = kev0.elapsed_time(kev1) / kreps
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), kernel_ms, m, clocks


def single_rank_reference(mcm, run: C2Runner, n_local, n_total, world):
    """The same global grid reduced WITHOUT the cross-rank path: rank 0 evaluates every rank's slice with the
    non-sharded kernel and merges the partials on the host (mcm_moments_merge, rank order)."""
    L, lib = mcm._lib, mcm._lib.lib()
    acc = None
    y = mcm.particles._empty(n_local, L.F64)
    t = run.t10
    for r in range(world):
        a = mcm.systematic_sample(n_total, mcm.Normal(np.pi, 0.1), permute=False, i0=r * n_local, n_total=n_total,
                                  n_local=n_local, as_particles=True)
        st = np.zeros(6)
        ins = (C.c_void_p * 1)(a.ptr); outs = (C.c_void_p * 1)(y.ptr)
        L.check(lib.mcm_chain(L.F64, run.tape_c, len(t.ins), run.lb, run.le, 0, run.consts.ctypes.data_as(L._dp), len(t.consts),
                              ins, 1, outs, 1, n_local, t.out_slots[0], st.ctypes.data_as(L._dp)))
        if acc is None:
            acc = st.copy()
        else:
            out = np.zeros(6)
            L.check(lib.mcm_moments_merge(acc.ctypes.data_as(L._dp), st.ctypes.data_as(L._dp), out.ctypes.data_as(L._dp)))
            acc = out
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="particles per GPU (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-families", action="store_true")
    ap.add_argument("--no-pcov", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chain-engine", type=int, default=0, help="mcm_chain_force_engine: 0 automatic, 1 registered (ahead of time), 2 interpreter, 3 generated")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return print_reference_line(args)

    import torch
    import torch.distributed as dist
    import mcm_b200 as mcm
    L = mcm._lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mcm.sharded.init_distributed(rank, world, local_rank)
    lib = L.lib()
    stream = torch.cuda.Stream()
    L.check(lib.mcm_set_stream(C.c_void_p(stream.cuda_stream)))
    L.check(lib.mcm_chain_force_engine(args.chain_engine))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the contract line: weak scaling (N per GPU fixed) unless --scaling strong -------------------------------
    strong = args.scaling == "strong"
    if strong:
        n_total = args.n
        lo, hi = mcm.sharded.shard_range(n_total, rank, world)
        N = hi - lo
    else:
        N, n_total, lo = args.n, world * args.n, rank * args.n
    run = C2Runner(mcm, N, n_total, lo, world, stream)
    elapsed_ms, kernel_ms, m, clocks = timed_c2(run, args.steps, args.warmup, barrier, torch, dist, world, local_rank)
    assert lib.mcm_chain_last_engine() in (1, 3), "chain10 must run on a straight-line (specialised) kernel"
    engine = {1: "registered (ahead of time)", 2: "interpreter", 3: "runtime-specialised (NVRTC)"}.get(lib.mcm_chain_last_engine())
    value = OPS_PER_PARTICLE * n_total * args.steps / (elapsed_ms * 1e-3) / 1e9

    # ---- parity of the cross-rank path, checked where the driver runs N > 1 ---------------------------------------
    parity = None
    if world > 1:
        detail = {}
        ok = True
        if rank == 0 and not strong:
            with torch.cuda.stream(stream):
                ref = single_rank_reference(mcm, run, N, n_total, world)
            got = np.array([m.n, m.mean, m.m2, m.min, m.max, m.sum])
            detail = {"n": [got[0], ref[0]], "rel_err_mean": abs(got[1] - ref[1]) / abs(ref[1]), "rel_err_m2": abs(got[2] - ref[2]) / ref[2],
                      "min_equal": bool(got[3] == ref[3]), "max_equal": bool(got[4] == ref[4]),
                      "how": "rank 0 re-evaluated all %d slices with the single-GPU kernel and merged them on the host" % world}
            ok = got[0] == ref[0] and detail["rel_err_mean"] <= 1e-12 and detail["rel_err_m2"] <= 1e-12 and detail["min_equal"] and detail["max_equal"]
        # every rank must hold the same bits
        bits = torch.tensor(np.array([m.n, m.mean, m.m2, m.min, m.max, m.sum]).view(np.int64), device="cuda")
        lo_b, hi_b = bits.clone(), bits.clone()
        dist.all_reduce(lo_b, op=dist.ReduceOp.MIN); dist.all_reduce(hi_b, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo_b, hi_b))
        detail["identical_on_all_ranks"] = same
        parity = {"moments": "ok" if (ok and same) else "FAILED", "detail": detail}
        barrier()

    # ---- end-to-end through the C ABI with HOST buffers (pinned input, H2D inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        e2e = e2e_bench(mcm, L, lib, run, N, n_total, world, barrier, torch, dist)

    # ---- strong scaling beside it (north_star: 1e8 particles in total, N/g per GPU) -------------------------------
    strong_line = None
    if not strong and world > 1:
        g_lo, g_hi = mcm.sharded.shard_range(args.n, rank, world)
        del run
        srun = C2Runner(mcm, g_hi - g_lo, args.n, g_lo, world, stream)
        s_ms, s_kernel_ms, sm, _ = timed_c2(srun, args.steps, args.warmup, barrier, torch, dist, world, local_rank, sample_clocks=False)
        strong_line = {"n_global": args.n, "particles_per_gpu": g_hi - g_lo, "ms_per_step": s_ms / args.steps,
                       "value": OPS_PER_PARTICLE * args.n * args.steps / (s_ms * 1e-3) / 1e9, "unit": "Gparticle-ops/s",
                       "kernel_ms": s_kernel_ms, "check": {"pmean": sm.mean, "pstd": sm.std(), "n": sm.n},
                       "limit": "per-step critical path = kernel on N/g particles + in-kernel peer exchange (P2P stores + flag wait, ~10 us) "
                                "+ event wait on the host; no NCCL launch, no memcpy"}
        run = srun
    elif not strong:
        strong_line = {"n_global": args.n, "note": "one GPU: identical to the weak-scaling line"}

    peak, peak_src = measured_peaks()
    achieved = ALGO_BYTES_PER_PARTICLE * N / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("chain10_f64_bytes_per_launch_1e8")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "chain10 straight-line kernel (Float64, fused statistics)",
                "engine": engine, "kernel_ms": kernel_ms,
                "note": "algorithmic 16 B/particle; this chain is FP64-pipe-bound (see fp64_pipe and DESIGN.md), the HBM-bound family is in hbm_family"}
    # the pipe that actually limits chain10: FP64 warp instructions per particle counted from the SASS by ncu
    # (profiles/: FP64 pipe issue slots x duration) against the measured issue rate of the pipe
    # (scripts/micro/fp64_rate.cu: one warp instruction per 2 cycles per SM sub-partition)
    fpath = os.path.join(ROOT, "profiles", "traffic.json")
    fp64_per_particle = 90.9
    if os.path.exists(fpath):
        fp64_per_particle = json.load(open(fpath)).get("chain10_fp64_inst_per_particle", fp64_per_particle)
    sm_hz = 1.965e9
    fp64_peak = 148 * 4 * 16 * sm_hz                      # FP64 lanes per second
    roofline["fp64_pipe"] = {"inst_per_particle": fp64_per_particle, "achieved_lanes_per_s": fp64_per_particle * N / (kernel_ms * 1e-3),
                             "peak_lanes_per_s": fp64_peak, "frac": fp64_per_particle * N / (kernel_ms * 1e-3) / fp64_peak,
                             "peak_source": "148 SMs x 4 sub-partitions x 16 lanes x 1.965 GHz (measured issue rate, scripts/micro/fp64_rate.cu)"}

    line = {
        "metric": "Gparticle-ops/s", "value": value, "unit": "Gparticle-ops/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: single uncertain scalar, N=1e8 Float64 particles %s, fused 10-primitive chain "
                               "(sin,^2,*,neg,exp,+,abs,sqrt,^1.5,/) + pmean/pstd in one launch" % ("in total" if strong else "per GPU"),
                   "particles_per_gpu": N, "particles_total": n_total,
                   "l2": "inputs and outputs (800 MB each per 1e8 particles) far exceed the 126 MB L2",
                   "parallelism": f"particle index sharded over {world} GPU(s); partial moments exchanged inside the kernel "
                                  f"over peer memory (P2P stores), merged in rank order; result in mapped pinned host memory",
                   "pipelining": "the 48-byte result of step k is collected while kernel k+1 runs; all results are collected inside the timed region"},
        "roofline": roofline, "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * world,
        "check": {"pmean": m.mean, "pstd": m.std(), "n": m.n},
    }
    if parity is not None:
        line["parity_check"] = parity
    if strong_line is not None:
        line["strong_scaling"] = strong_line
    del run

    if not args.no_pcov:
        pc = pcov_bench(mcm, L, lib, stream, rank, world)
        if rank == 0:
            line["pcov"] = pc
            if parity is not None and "parity" in pc:
                parity["pcov"] = pc["parity"]["status"]
    if parity is not None:
        line["parity_check"]["status"] = "ok" if all(v == "ok" for k, v in parity.items() if k in ("moments", "pcov")) else "FAILED"
    if rank == 0 and world == 1:
        if not args.no_families:
            line["hbm_family"] = hbm_family(mcm, L, lib, stream, N, peak)
        if not args.no_c3:
            line["c3_pendulum"] = c3_bench(mcm, L, lib, stream, peak)
        if not args.no_c5:
            line["c5_pipeline"] = c5_bench(mcm, L, lib, stream)
        if not args.no_cpu_baseline:
            r = cpu_reference_arm(steps=5, warmup=1)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": "Gparticle-ops/s", "cores": r["threads"], "kind": "port",
                "sample": f"chain10 + pmean/pstd on N={r['n']} Float64 (the same N as the CUDA arm), mean of 5 passes after 1 warm-up, "
                          f"SLEEF 3.8 SIMD loop-per-primitive port (oracle/ref_cpu.cpp), persistent pool of {r['threads']} host threads",
                "check": {"pmean": r["mean"], "pstd": r["std"]}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_bench(mcm, L, lib, run, N, n_total, world, barrier, torch, dist):
    """The same step through the reference-facing call with HOST buffers: mcm_chain_host copies pinned host input
    to the device in chunks overlapped with the fused kernel and returns the six moments; for N > 1 ranks the
    per-rank moments are merged through the library's communicator."""
    t10 = run.t10
    hp = C.c_void_p()
    wc = os.environ.get("MCM_E2E_WC", "0") == "1"         # A/B: write-combined pinned input (no snoop on the DMA reads)
    L.check((lib.mcm_host_alloc_wc if wc else lib.mcm_host_alloc)(C.byref(hp), N * 8))
    L.check(lib.mcm_download(hp, C.c_void_p(run.a.ptr), N * 8))
    hin = (C.c_void_p * 1)(hp.value)
    e2e_steps = 5
    st = np.zeros(6)

    def e2e_step():
        L.check(lib.mcm_chain_host(L.F64, run.tape_c, len(t10.ins), run.lb, run.le, 0, run.consts.ctypes.data_as(L._dp), len(t10.consts),
                                   hin, 1, None, 1, N, t10.out_slots[0], st.ctypes.data_as(L._dp)))
        mm = mcm.Moments(*st)
        return mcm.sharded.allmerge(mm) if world > 1 else mm
    for _ in range(2):
        me = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        me = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = OPS_PER_PARTICLE * n_total * e2e_steps / float(te.item()) / 1e9
    out = {"value": e2e_val, "unit": "Gparticle-ops/s", "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 48,
           "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
           "aggregate_h2d_GBps": 8 * n_total / (float(te.item()) / e2e_steps) / 1e9,
           "how": "mcm_chain_host: pinned host input -> chunked H2D overlapped with the fused kernel -> 6 moments D2H; y not returned",
           "host_buffer": "write-combined pinned" if wc else "pinned",
           "check": {"pmean": me.mean, "pstd": me.std()}}
    L.check(lib.mcm_host_free(hp))
    return out


def pcov_bench(mcm, L, lib, stream, rank, world):
    """Config C4: pcov of M = 4096 uncertain variables x N = 1e5 Float64 particles per GPU (the particle index is
    sharded: global N = world * 1e5).  TFLOP/s in the GEMM convention 2*M^2*N (the symmetric half is computed).
    Engines: tcgen05/TMEM int8 tensor cores with an Ozaki split (default) and the FP64 tensor pipe (DMMA)."""
    import torch
    M, n = 4096, 100_000

    def build(r, Mv):
        base = [mcm.systematic_sample(n, mcm.Normal(0.0, 1.0), permute=True, seed=1 + j + 64 * r, as_particles=True)
                for j in range(64)]
        return [mcm.lincomb([1.0 + 0.001 * j, 0.5], [base[j % 64], base[(j * 7 + 3) % 64]]) for j in range(Mv)]
    v = build(rank, M)
    ptrs = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in v])
    dC = mcm.particles._alloc(M * M * 8)
    out = {"config": f"C4: M={M} variables x N={n} Float64 particles per GPU, corrected covariance, {world} GPU(s)",
           "flops_convention": "2*M^2*N_global (GEMM convention; only the upper triangle of tiles is computed)"}
    with torch.cuda.stream(stream):
        for name, mode in (("tcgen05_int8_ozaki", 1), ("fp64_dmma", 0)):
            lib.mcm_gram_use_tensor(mode)

            def run():
                if world > 1:
                    L.check(lib.mcm_gram_sharded(L.F64, ptrs, M, n, 1, C.c_void_p(dC)))
                else:
                    L.check(lib.mcm_gram(L.F64, ptrs, M, n, 1, None, C.c_void_p(dC), None))
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record(stream)
            for _ in range(reps):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            out[name] = {"ms": ms, "TFLOP/s": 2.0 * M * M * n * world / (ms * 1e-3) / 1e12,
                         "engine_used": "tcgen05" if lib.mcm_gram_last_engine() == 1 else "dmma"}
        lib.mcm_gram_use_tensor(1)
        if world == 1:
            # heavy-tailed rows (the outputs of nonlinear chains: exp of a Gaussian) stay on the tensor cores with 7 / 8 digits
            vh = [mcm.exp(1.5 * p) for p in v[:64]]
            vh = [mcm.lincomb([1.0 + 0.001 * j, 0.5], [vh[j % 64], vh[(j * 7 + 3) % 64]]) for j in range(M)]
            ph = (C.c_void_p * M)(*[C.c_void_p(p.ptr) for p in vh])
            L.check(lib.mcm_gram(L.F64, ph, M, n, 1, None, C.c_void_p(dC), None)); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                L.check(lib.mcm_gram(L.F64, ph, M, n, 1, None, C.c_void_p(dC), None))
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out["heavy_tailed_rows"] = {"data": "lognormal rows exp(1.5 z) (max/sigma ~ 50-100)", "ms": ms,
                                        "TFLOP/s": 2.0 * M * M * n / (ms * 1e-3) / 1e12,
                                        "engine_used": "tcgen05" if lib.mcm_gram_last_engine() == 1 else "dmma",
                                        "digits": int(lib.mcm_gram_last_digits())}
            del vh
        if world > 1:
            # parity of the sharded pcov: the leading Ms x Ms block against a single-rank Gram of the concatenated slices
            import torch.distributed as dist
            Ms = 128
            L.check(lib.mcm_gram_sharded(L.F64, ptrs, M, n, 1, C.c_void_p(dC)))
            Cs = np.empty((M, M))
            L.check(lib.mcm_download(Cs.ctypes.data_as(C.c_void_p), C.c_void_p(dC), Cs.nbytes))
            if rank == 0:
                cat = [mcm.particles._empty(n * world, L.F64) for _ in range(Ms)]
                for r in range(world):
                    vr = build(r, Ms)
                    for j in range(Ms):
                        L.check(lib.mcm_copy(C.c_void_p(cat[j].ptr + 8 * n * r), C.c_void_p(vr[j].ptr), 8 * n))
                ref = mcm.pcov(cat)
                err = float(np.max(np.abs(Cs[:Ms, :Ms] - ref)) / np.max(np.abs(ref)))
                out["parity"] = {"status": "ok" if err <= 1e-12 else "FAILED", "max_rel_err": err,
                                 "how": f"leading {Ms}x{Ms} block of the sharded pcov vs mcm_gram on the concatenated N={n * world} particles on rank 0"}
            dist.barrier()
    L.check(lib.mcm_free(C.c_void_p(dC)))
    return out


def c3_bench(mcm, L, lib, stream, peak):
    """Config C3: RK4 damped pendulum, N = 1e7 particles x 1000 steps, Float64, five uncertain inputs
    (theta0, omega0, g, L, c).  (1) the whole run fused in registers: ONE launch, 56 B/particle in total;
    (2) trajectory-saving variant: one launch per step, the state written back every step as `sol.u` does
    (16 B state + 24 B parameters read, 16 B written per particle-step)."""
    import torch
    from mcm_b200 import chains, tape as tp
    n, steps = 10_000_000, 1000
    ins = [mcm.systematic_sample(n, d, permute=True, seed=11 + k, as_particles=True)
           for k, d in enumerate((mcm.Normal(0.0, 1e-12), mcm.Normal(np.pi / 3, 0.02), mcm.Normal(9.79, 0.02),
                                  mcm.Normal(1.0, 0.01), mcm.Normal(0.1, 0.01)))]
    t = chains.pendulum_rk4_tape(0.01)
    o = [mcm.particles._empty(n, L.F64) for _ in range(2)]
    out = {"config": f"C3: RK4 damped pendulum, N={n} Float64 particles x {steps} steps, 5 uncertain inputs, dt=0.01"}
    with torch.cuda.stream(stream):
        st = tp.run_tape(t, [p.ptr for p in ins], n, L.F64, [o[0].ptr, o[1].ptr], stats_output=0, reps=10)   # warm-up / compile
        best = None
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            st = tp.run_tape(t, [p.ptr for p in ins], n, L.F64, [o[0].ptr, o[1].ptr], stats_output=0, reps=steps)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        out["fused_in_registers"] = {"ms": best, "Gparticle_steps_per_s": n * steps / (best * 1e-3) / 1e9,
                                     "Gparticle_ops_per_s": 44 * n * steps / (best * 1e-3) / 1e9,
                                     "bound": "FP64 pipe (4 sin + ~40 flops per particle-step; 56 B/particle for the whole run)",
                                     "engine": lib.mcm_chain_last_engine(), "check": {"pmean_theta": st[1], "pstd_theta": float(np.sqrt(st[2] / (st[0] - 1)))}}
        # trajectory-saving: ping-pong state buffers over a ring of 8 output pairs (each launch reads the previous
        # state and the parameters from HBM and writes the new state to a DIFFERENT buffer: the traffic of sol.u)
        ring = [[mcm.particles._empty(n, L.F64) for _ in range(2)] for _ in range(8)]
        cur = [ins[0], ins[1]]
        for k in range(8):
            tp.run_tape(t, [cur[0].ptr, cur[1].ptr, ins[2].ptr, ins[3].ptr, ins[4].ptr], n, L.F64, [ring[k][0].ptr, ring[k][1].ptr], reps=1)
            cur = ring[k]
        cur = [ins[0], ins[1]]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            nxt = ring[k % 8]
            tp.run_tape(t, [cur[0].ptr, cur[1].ptr, ins[2].ptr, ins[3].ptr, ins[4].ptr], n, L.F64, [nxt[0].ptr, nxt[1].ptr], reps=1)
            cur = nxt
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        mom = mcm.moments(cur[0])
        gbs = 56.0 * n * steps / (ms * 1e-3) / 1e9
        out["trajectory_saved"] = {"ms": ms, "ms_per_step": ms / steps, "Gparticle_steps_per_s": n * steps / (ms * 1e-3) / 1e9,
                                   "bytes_per_particle_step": 56, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak,
                                   "launches": steps, "check": {"pmean_theta": mom.mean, "pstd_theta": mom.std()}}
    return out


def c5_bench(mcm, L, lib, stream):
    """Config C5 on one GPU: d = 1024 uncertain variables x N = 1e6 Float64 particles: build (keyed systematic
    columns) -> transform_moments to (m, S) -> bootstrap with one shared index vector -> pquantile of every variable
    at {0.025, 0.5, 0.975}.  Wall clock around synchronising public-API calls (they include the host work:
    two 1024^2 Cholesky factorisations + a triangular solve inside transform_moments), best of 2."""
    import torch
    d, n = 1024, 1_000_000
    rng = np.random.default_rng(0)
    Aq = rng.normal(size=(d, d))
    S = Aq @ Aq.T / d + np.eye(d)
    m = np.linspace(-1, 1, d)
    q = np.array([0.025, 0.5, 0.975])

    def timed(fn, reps=2):
        best, res = None, None
        for _ in range(reps + 1):          # first pass warms the pool allocator
            torch.cuda.synchronize(); t = time.perf_counter(); res = fn(); mcm.sync(); dt = time.perf_counter() - t
            best = dt if best is None or _ == 1 else min(best, dt)
        return best * 1e3, res
    out = {"config": f"C5: d={d} variables x N={n} Float64 particles, one GPU"}
    with torch.cuda.stream(stream):
        out["build_ms"], cols = timed(lambda: mcm.systematic_batch(n, d, permute=True, seed=1))
        out["transform_moments_ms"], Y = timed(lambda: mcm.transform_moments(cols, m, S))
        out["apply_engine"] = "tcgen05" if lib.mcm_pgemm_last_engine() == 1 else "dmma"
        del cols
        out["bootstrap_ms"], R = timed(lambda: mcm.bootstrap(Y, n, seed=3))
        out["pquantile_ms"], Q = timed(lambda: mcm.pquantile(R, q))
        out["total_ms"] = out["build_ms"] + out["transform_moments_ms"] + out["bootstrap_ms"] + out["pquantile_ms"]
        out["check"] = {"max_abs_mean_err": float(np.max(np.abs(np.array(mcm.pmean(Y)) - m))),
                        "max_abs_cov_err": float(np.max(np.abs(mcm.pcov(Y) - S))),
                        "q_of_var0": [float(v) for v in Q[0]]}
    return out


def hbm_family(mcm, L, lib, stream, N, peak):
    """The low-intensity chains the >= 80 % HBM target is about: CUDA-event timing, inputs > L2."""
    import torch
    out = {}
    a = mcm.systematic_sample(N, mcm.Normal(np.pi, 0.1), permute=False, as_particles=True)
    y = mcm.particles._empty(N, L.F64)
    b = mcm.particles._empty(N, L.F64); c = mcm.particles._empty(N, L.F64)
    L.check(lib.mcm_affine(L.F64, C.c_void_p(a.ptr), 0.5, 1.0, C.c_void_p(b.ptr), N))
    L.check(lib.mcm_affine(L.F64, C.c_void_p(a.ptr), -0.25, 2.0, C.c_void_p(c.ptr), N))
    coeffs = np.array([0.5, -1.0, 0.25, 2.0, -0.125, 0.3, 0.01, -0.02, 0.003])
    ptrs = (C.c_void_p * 3)(a.ptr, b.ptr, c.ptr)
    cs = np.array([1.0, 0.01 / 6, 0.02 / 6])
    mom = np.zeros(6)
    cases = {
        "affine (alpha*a+beta)": (16, lambda: lib.mcm_affine(L.F64, C.c_void_p(a.ptr), 1.5, 0.5, C.c_void_p(y.ptr), N)),
        "poly8 (Horner, 8 FMA)": (16, lambda: lib.mcm_polyval(L.F64, C.c_void_p(a.ptr), coeffs.ctypes.data_as(L._dp), 9, 1, C.c_void_p(y.ptr), N)),
        "axpy3 (u + h*k1 + h*k2, RK4 stage)": (32, lambda: lib.mcm_lincomb(L.F64, 3, ptrs, cs.ctypes.data_as(L._dp), C.c_void_p(y.ptr), N)),
        "add (p1 + p2)": (24, lambda: lib.mcm_binary(L.ENUMS["MCM_B_ADD"], L.F64, C.c_void_p(a.ptr), 0.0, C.c_void_p(b.ptr), 0.0, C.c_void_p(y.ptr), N)),
        "sqrt": (16, lambda: lib.mcm_unary(L.ENUMS["MCM_U_SQRT"], L.F64, C.c_void_p(a.ptr), C.c_void_p(y.ptr), N)),
        "exp": (16, lambda: lib.mcm_unary(L.ENUMS["MCM_U_EXP"], L.F64, C.c_void_p(a.ptr), C.c_void_p(y.ptr), N)),
        "sin": (16, lambda: lib.mcm_unary(L.ENUMS["MCM_U_SIN"], L.F64, C.c_void_p(a.ptr), C.c_void_p(y.ptr), N)),
        "pmean+pstd (one pass)": (8, lambda: lib.mcm_moments(L.F64, C.c_void_p(a.ptr), N, mom.ctypes.data_as(L._dp))),
        "pmean+pstd kernel only (mcm_moments_enqueue: result left on device, no host round trip per call)":
            (8, lambda: lib.mcm_moments_enqueue(L.F64, C.c_void_p(a.ptr), N)),
    }
    with torch.cuda.stream(stream):
        for name, (bpp, fn) in cases.items():
            for _ in range(3):
                L.check(fn())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record(stream)
            for _ in range(reps):
                L.check(fn())
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = bpp * N / (ms * 1e-3) / 1e9
            out[name] = {"ms": ms, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak, "bytes_per_particle": bpp}
    return out


if __name__ == "__main__":
    main()
