#!/usr/bin/env python
"""Top stall-sample instructions of one kernel from `ncu -i rep --page source --csv` output.

    ncu -i x.ncu-rep --page source --csv > /tmp/src.csv ; python tools/ncu_hot.py /tmp/src.csv <kernel-section-index> [top]
"""
import csv, sys
path, sec, top = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
s, e = starts[sec], starts[sec + 1]
print(rows[s][1])
hdr = rows[s + 1]
ix = {n: i for i, n in enumerate(hdr)}
body = rows[s + 2:e]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ix[n]] or 0) for r in body) for n in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = {n[6:]: int(r[ix[n]]) for n in stalls if int(r[ix[n]] or 0) > 0}
    print(f"{i:5d} {r[ix['# Samples']]:>7} exec={r[ix['Instructions Executed']]:>10} {r[ix['Source']].strip()[:70]:70s} {st}")
