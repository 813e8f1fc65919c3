"""Summarise an ncu report of one kernel: headline metrics, stall sampling totals, the instructions with the most
stall samples.   python scripts/ncu_hot.py gpurun_out/x.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sass__thread_inst_executed_true_per_opcode", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active"]
for k in want:
    if k in d:
        print("%-75s %s" % (k, d[k]))
tot = float(d.get("smsp__pcsamp_sample_count", 0) or 0)
st = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and v not in ("", "n/a")), reverse=True)
print("stall samples (all): total", tot)
for v, k in st[:10]:
    print("   %6.1f%%  %s" % (100 * v / max(tot, 1), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
data = rows[2:]
tots = sum(int(r[ix["# Samples"]]) for r in data)
ie = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("warp instructions executed", ie, " samples", tots)
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    why = sorted(((int(r[ix[c]] or 0), c[6:]) for c in cols), reverse=True)[:2]
    print("%5.2f%%  x%-9s %-52s %s" % (100 * int(r[ix["# Samples"]]) / tots, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:52],
                                        " ".join("%s=%d" % (b, a) for a, b in why if a)))
# opcode mix, dynamic
mix = {}
for r in data:
    op = r[ix["Source"]].strip().split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[ix["Instructions Executed"]])
print("dynamic opcode mix (warp instructions, % of all):")
print("  ", ", ".join("%s %.1f%%" % (k, 100 * v / ie) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:22]))
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
print("FP64 warp instructions %.0f (%.1f%% of all)" % (fp64, 100 * fp64 / ie))
