"""Stall samples, shared-memory wavefronts and executed warp instructions per SOURCE LINE of an ncu report
(kernel built with -lineinfo, captured with --import-source on).   python scripts/ncu_lines.py x.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[2]
ix = {}
for i, k in enumerate(h):
    ix.setdefault(k, i)
lines = [r for r in rows[3:] if len(r) == len(h) and r[ix["Line No"]].isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in lines) or 1
wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in lines) or 1
ie = sum(int(r[ix["Instructions Executed"]] or 0) for r in lines) or 1
print("samples %d  shared wavefronts %d (ideal %d)  warp instructions %d" % (tot, wf, sum(int(r[ix["L1 Wavefronts Shared Ideal"]] or 0) for r in lines), ie))
for r in sorted(lines, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
    print("%5.1f%% samples %5.1f%% wavefronts %5.1f%% inst  L%-4s %s" % (
        100 * int(r[ix["# Samples"]] or 0) / tot, 100 * int(r[ix["L1 Wavefronts Shared"]] or 0) / wf,
        100 * int(r[ix["Instructions Executed"]] or 0) / ie, r[ix["Line No"]], r[1].strip()[:110]))
