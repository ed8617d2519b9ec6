"""Per-kernel SASS mnemonic table of the built library -> profiles/<name>.md (no GPU needed).

    python tools/sass_table.py [profiles/r02_sass_mnemonics.md]

Blackwell-native evidence (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA -> UTMALDG, tcgen05.commit -> UTCBAR; HMMA would mean the legacy mma.sync path.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "two-tower-model-v2_b200" / "lib" / "libtt_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UTCATOM", "SYNCS", "HMMA", "FFMA", "FFMA2", "LDG", "STG",
        "LDS", "STS", "SHFL", "MUFU", "ACQBULK", "ATOM", "RED"]


def main():
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r02_sass_mnemonics.md"
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            kernels[cur][base] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS mnemonic counts per kernel (libtt_b200.so, sm_100a)", "",
             "`cuobjdump -sass two-tower-model-v2_b200/lib/libtt_b200.so`, instruction counts per kernel (static, not executed).",
             "UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, UTCBAR = tcgen05.commit; no kernel uses HMMA.", "",
             "| kernel | instrs | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for (name, c), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r"\(.*", "", pretty).replace("void ", "")
        lines.append(f"| `{pretty[:70]}` | {c['_total']} | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")
    out.write_text("\n".join(lines) + "\n")
    print(f"wrote {out} ({len(kernels)} kernels)")


if __name__ == "__main__":
    main()
