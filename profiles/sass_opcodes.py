#!/usr/bin/env python
"""SASS opcode census of the shipped library, per kernel: `python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt`.
What to read: DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64; there is no tcgen05 kind for FP64), LDGSTS = cp.async,
UTMALDG / UBLKCP = TMA, UTC*MMA / LDTM = tcgen05 (none: FP64 path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stab_b200", "libstabgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDGSTS", "UTMALDG", "UBLKCP", "UTCMMA", "LDTM", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "LDL", "STL"]
rows = []
name, cnt, total = None, None, 0
archs = set(re.findall(r"arch = (sm_\w+)", txt))
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, cnt, total))
        name, cnt, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        total += 1
        op = m.group(1)
        for c in cols:
            if op.startswith(c):
                cnt[c] += 1
                break
if name:
    rows.append((name, cnt, total))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: architectures {sorted(archs)}; {len(rows)} kernels")
print("# " + " ".join(f"{c:>7s}" for c in ["instr"] + cols) + "  kernel")
tot = collections.Counter()
for (nm, cnt, total), d in sorted(zip(rows, dem), key=lambda x: -x[0][2]):
    short = re.sub(r"\(.*", "", d).replace("stab::", "").replace("void ", "")
    print("  " + " ".join(f"{v:7d}" for v in [total] + [cnt[c] for c in cols]) + "  " + short)
    tot.update(cnt); tot["instr"] += total
print("  " + " ".join(f"{v:7d}" for v in [tot["instr"]] + [tot[c] for c in cols]) + "  TOTAL")
