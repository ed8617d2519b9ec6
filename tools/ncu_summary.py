#!/usr/bin/env python
"""Summarises `ncu --set full` reports (.ncu-rep) into a markdown table (one row per captured launch) and
collects the DRAM traffic per launch of the main scan kernel into profiles/scan_traffic.json, which bench.py
reports as roofline.traffic.

    python tools/ncu_summary.py --tag r01 gpurun_out/scan_nq4096.ncu-rep [...]  > profiles/r01_ncu_summary.md
"""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc smem %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_elapsed.max", "cycles"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def short(name):
    m = re.search(r"(flat_scan_kernel<[^>]*>|[A-Za-z0-9_]+)\s*\(", name)
    return m.group(1) if m else name[:40]


def main():
    args = sys.argv[1:]
    tag = "r01"
    if "--tag" in args:
        i = args.index("--tag")
        tag = args[i + 1]
        del args[i:i + 2]
    traffic_path = ROOT / "profiles" / "scan_traffic.json"
    traffic = json.loads(traffic_path.read_text()) if traffic_path.exists() else {}
    print(f"# ncu --set full summaries ({tag})\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` on one B200 (see tools/gpu_r0N_evidence.sh);")
    print("per-launch values, cold cache, serialised; durations are NOT bench values.\n")
    print("| report | kernel | grid | " + " | ".join(n for _, n in METRICS) + " |")
    print("|---|---|---|" + "---|" * len(METRICS))
    for rep in args:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            cells = []
            for m, _ in METRICS:
                if m in ix:
                    v, u = r[ix[m]], units[ix[m]]
                    try:
                        cells.append(f"{float(v):.4g} {u}".strip())
                    except ValueError:
                        cells.append(v)
                else:
                    cells.append("-")
            k = short(r[ix["Kernel Name"]])
            print(f"| {Path(rep).stem} | `{k}` | {r[ix['Grid Size']]} | " + " | ".join(cells) + " |")
            m = re.search(r"scan_nq(\d+)", Path(rep).stem)
            if m and re.search(r"flat_scan_kernel<\(int\)\d+, \(bool\)0|flat_scan_kernel<\d+, 0", r[ix["Kernel Name"]]):
                rd = float(r[ix["dram__bytes_read.sum"]]) * UNIT.get(units[ix["dram__bytes_read.sum"]], 1.0)
                wr = float(r[ix["dram__bytes_write.sum"]]) * UNIT.get(units[ix["dram__bytes_write.sum"]], 1.0)
                traffic[f"10000000x384_nq{m.group(1)}"] = rd + wr        # the evidence scripts capture the headline catalog
    traffic_path.write_text(json.dumps(traffic, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
