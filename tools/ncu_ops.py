#!/usr/bin/env python
"""Per-opcode stall-sample / instruction / shared-wavefront table from an .ncu-rep (source page).
usage: ncu_ops.py report.ncu-rep [kernel-name-regex]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2]
tot = sum(int(r[ix["# Samples"]]) for r in data)
g = defaultdict(lambda: [0, 0, 0, 0, 0])
for r in data:
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    parts = op.split(".")
    key = parts[0] + ("." + ".".join(parts[1:3]) if parts[0] in ("LDS", "STS", "LDG", "STG", "LDGSTS", "BAR") else "")
    g[key][0] += int(r[ix["# Samples"]]); g[key][1] += int(r[ix["Instructions Executed"]])
    for j, c in ((2, "L1 Wavefronts Shared"), (3, "L1 Wavefronts Shared Ideal"), (4, "L2 Theoretical Sectors Global")):
        try: g[key][j] += int(r[ix[c]])
        except Exception: pass
print("total samples", tot)
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{k:24s} samples {v[0]:7d} ({100 * v[0] / tot:5.1f}%)  inst {v[1]:10d}  smem wf {v[2]:10d} ideal {v[3]:10d}  L2 sectors {v[4]}")
