"""Top warp-stall locations of a kernel from an .ncu-rep (SASS view): python tools/ncu_src_top.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        c_s, c_src, c_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
        j, data = i + 2, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            if len(r) > c_s:
                try:
                    data.append((float(r[c_s]), j - i - 2, r[c_src].strip(), r[c_ex]))
                except ValueError:
                    pass
            j += 1
        tot = sum(d[0] for d in data) or 1
        print(f"== {name[:100]}  samples {tot:.0f}, instructions {len(data)}")
        for v, idx, src, ex in sorted(data, reverse=True)[:n]:
            print(f"{v / tot * 100:5.1f}%  #{idx:5d} x{ex:>8}  {src[:100]}")
        i = j
    else:
        i += 1
