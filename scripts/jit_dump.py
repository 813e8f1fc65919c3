"""Generate + compile (NVRTC, no GPU needed) the kernel of a canonical chain and summarise its SASS:
    python scripts/jit_dump.py [chain10|readme|rk4] [f64|f32] [nostats]
Writes /tmp/mcm_jit/<name>.{cu,cubin,sass} and prints the opcode histogram of the hot loop."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mcm_b200 as mcm
from mcm_b200 import chains, tape as tp

L = mcm._lib
lib = L.load()
for a_ in sys.argv[1:]:
    if a_.startswith("opt"):
        w_, v_ = a_[3:].split("=")
        L.check(lib.mcm_chain_jit_option(int(w_), int(v_)))
name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("opt") else "chain10"
f64 = "f32" not in sys.argv
stats = "nostats" not in sys.argv
t = {"chain10": lambda: tp.trace(chains.chain10, 1), "readme": lambda: tp.trace(chains.readme_expr, 1),
     "rk4": lambda: chains.pendulum_rk4_tape(0.01)}[name]()
os.makedirs("/tmp/mcm_jit", exist_ok=True)
base = "/tmp/mcm_jit/%s_%s%s" % (name, "f64" if f64 else "f32", "" if stats else "_nostats")
src = C.create_string_buffer(1 << 18)
nb = C.c_size_t(0)
consts = np.asarray(t.consts if t.consts else [0.0])
lb, le = t.bounds
rc = lib.mcm_chain_jit_source(L.F64 if f64 else L.F32, t.as_ctypes(), len(t.ins), lb, le, consts.ctypes.data_as(L._dp), len(t.consts),
                              t.out_slots[0] if stats else -1, 1, src, len(src), C.byref(nb), (base + ".cubin").encode())
open(base + ".cu", "w").write(src.value.decode())
if rc:
    print("FAILED:", lib.mcm_last_error().decode()); sys.exit(1)
print("cubin bytes", nb.value)
sass = subprocess.run(["cuobjdump", "-sass", base + ".cubin"], capture_output=True, text=True).stdout
open(base + ".sass", "w").write(sass)
res = subprocess.run(["cuobjdump", "-res-usage", base + ".cubin"], capture_output=True, text=True).stdout
print([l for l in res.splitlines() if "REG" in l])
# main kernel only
k = sass[sass.find("Function : mcm_jit_kernel"):]
nxt = k.find("Function :", 10)
k = k[:nxt] if nxt > 0 else k
ins = re.findall(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", k, flags=re.M)
print("kernel instructions (static):", len(ins))
# every loop (backward branch region) of more than 60 instructions: the fast body's loop comes first, the exact
# replay's copy (library fallbacks, CALLs) later
loops = []
for addr, text in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < int(addr, 16) and (int(addr, 16) - int(m.group(1), 16)) // 16 > 60:
        loops.append((int(m.group(1), 16), int(addr, 16)))
for best in loops:
    print("loop: 0x%x .. 0x%x = %d instructions (static)" % (best[0], best[1], (best[1] - best[0]) // 16 + 1))
    hist = {}
    for addr, text in ins:
        a = int(addr, 16)
        if best[0] <= a <= best[1]:
            op = re.sub(r"^@!?U?P\w+\s+", "", text.strip()).split()[0].split(".")[0]
            hist[op] = hist.get(op, 0) + 1
    fp64 = sum(v for k_, v in hist.items() if k_ in ("DFMA", "DADD", "DMUL", "DSETP"))
    print("  FP64 in loop:", fp64, " other:", sum(hist.values()) - fp64)
    print("  ", sorted(hist.items(), key=lambda kv: -kv[1]))
