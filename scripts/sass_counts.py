"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA = UTCIMMA / UTCHMMA / UTCQMMA, TMEM
loads LDTM, bulk TMA copies UBLKCP, tensor-map TMA UTMALDG, mbarrier SYNCS, FP64 tensor DMMA) in libmcm_b200.so:
    python scripts/sass_counts.py > profiles/r2_sass_counts.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "montecarlomeasurements.jl_b200", "libmcm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCCP", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "DMMA", "DFMA", "MUFU.RSQ64H", "MUFU.RCP64H", "REDUX", "VOTE"]
cur, counts = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = {}
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for p in pats:
            if op.startswith(p):
                counts[cur][p] = counts[cur].get(p, 0) + 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts.keys()), capture_output=True, text=True).stdout.splitlines()
tot = {}
print("%-100s %s" % ("kernel", "mnemonic counts (static)"))
for (k, c), name in sorted(zip(counts.items(), demangle), key=lambda kv: kv[1]):
    interesting = {p: v for p, v in c.items() if p not in ("DFMA", "VOTE", "REDUX")}
    for p, v in c.items():
        tot[p] = tot.get(p, 0) + v
    if interesting:
        print("%-100s %s" % (name[:100], " ".join("%s=%d" % kv for kv in sorted(c.items()))))
print("\nTOTAL over %d kernels: %s" % (len(counts), " ".join("%s=%d" % kv for kv in sorted(tot.items()))))
