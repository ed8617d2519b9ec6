#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python tools/summarise_launches.py gpurun_out/launches.csv [--skip N] > profiles/rNN_launches_<tag>.md

Kernel names are shortened to the function name (template arguments kept for our own kernels).
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = name.strip()
    if name.startswith("void "):
        name = name[5:]
    m = re.match(r"(tt::[A-Za-z0-9_]+(?:<[^>(]*>)?)", name)
    if m:
        return m.group(1)
    m = re.match(r"([A-Za-z0-9_:<>]+?)[<(]", name)
    base = m.group(1) if m else name[:60]
    if "distribution_elementwise" in name:
        return "at::distribution_elementwise (randn)"
    return base[:70]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
    rows = rows[skip:]
    agg = OrderedDict()
    for _, k, g, b, ns in rows:
        key = (k, g, b)
        a = agg.setdefault(key, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += ns; a[2] = min(a[2], ns); a[3] = max(a[3], ns)
    total = sum(a[1] for a in agg.values())
    print(f"source: {path}; launches: {len(rows)}; total kernel time {total/1e6:.3f} ms (cold-cache, serialised under ncu)\n")
    print("| kernel | grid | block | launches | avg us | min us | max us | share of total |")
    print("|---|---|---|---|---|---|---|---|")
    for (k, g, b), a in agg.items():
        print(f"| `{k}` | {g} | {b} | {a[0]} | {a[1]/a[0]/1e3:.1f} | {a[2]/1e3:.1f} | {a[3]/1e3:.1f} | {a[1]/total*100:.2f}% |")


if __name__ == "__main__":
    main()
