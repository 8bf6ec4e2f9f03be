#!/usr/bin/env python
"""Top-N instructions by stall samples from an `ncu --page source --csv` dump (plain or .gz), with the
dominant stall reasons and the source position in SASS order.  usage: ncu_top.py source.csv[.gz] [N]"""
import csv, gzip, sys
fn = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = (gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)).read()
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:topn]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:6d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{n}={v}" for v, n in st if v))

# optional: per-region stall breakdown  (usage: ... N lo1-hi1,lo2-hi2)
if len(sys.argv) > 3:
    for rg in sys.argv[3].split(","):
        lo, hi = [int(x) for x in rg.split("-")]
        agg = {}
        n = 0; inst = 0
        for r in data[lo:hi]:
            n += int(r[ix["# Samples"]]); inst += int(r[ix["Instructions Executed"]])
            for c in stall_cols:
                if "Not Issued" in c: continue
                agg[c[6:]] = agg.get(c[6:], 0) + int(r[ix[c]] or 0)
        print(f"region {lo}-{hi}: samples {n} ({100*n/tot:.1f}%), warp-instructions {inst}")
        print("   " + "  ".join(f"{k}={v}({100*v/max(n,1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
