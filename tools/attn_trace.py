"""Pipeline timeline of CTA 0 of the fused attention-pooling kernel (TT_B200_ATTN_TRACE): cycles relative to the tile's
TMA issue, per tile.  python tools/attn_trace.py [B]"""
import os
import sys
import tempfile
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
path = os.path.join(tempfile.gettempdir(), "attn_trace.txt")
os.environ["TT_B200_ATTN_TRACE"] = path
import two_tower_model_v2_b200 as pkg  # noqa: E402

B, S, D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 50, 384
g = torch.Generator(device="cuda").manual_seed(99)
x = torch.randn((B, S, D), device="cuda", generator=g)
w = torch.tensor([1.0, 5.0, 10.0], device="cuda")[torch.randint(0, 3, (B, S), device="cuda", generator=g)]
torch.manual_seed(0)
m = pkg.BuyerTower(D, "attention").cuda()
with torch.no_grad():
    for _ in range(3):
        m(x, w)
torch.cuda.synchronize()
rows = [[int(v) for v in l.split()] for l in open(path)]
base = min(v for r in rows for v in r if v > 0)
for t, r in enumerate(rows):
    if r[0] == 0:
        continue
    t0 = r[0]
    f = lambda i: r[i] - t0 if r[i] else None
    print(f"tile {t:2d} @ {t0 - base:7d}: " + " | ".join(
        f"kb{kb}: tma {f(kb)} raw {f(6 + kb)} split {f(12 + kb)} mma_sees {f(18 + kb)} commit {f(24 + kb)}" for kb in range(6))
        + f" | epi sees {f(30)} epi done {f(31)}")
