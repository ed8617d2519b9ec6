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
names = {1: "buf_free seen", 0: "tma issued", 2: "raw0", 3: "raw1", 4: "raw2", 5: "raw3", 6: "raw4", 7: "raw5",
         8: "split0", 9: "split1", 10: "split2", 11: "split3", 12: "split4", 13: "split5",
         14: "mma_sees0", 15: "mma_sees1", 16: "mma_sees2", 17: "mma_sees3", 18: "mma_sees4", 19: "mma_sees5",
         25: "acc_empty seen", 20: "commit issued", 21: "epi sees acc_full", 22: "logits written", 23: "pool sees logits", 24: "pool done"}
order = [1, 0, 2, 8, 14, 3, 9, 15, 4, 10, 16, 5, 11, 17, 6, 12, 18, 7, 13, 19, 25, 20, 21, 22, 23, 24]
for t, r in enumerate(rows):
    if r[0] == 0:
        continue
    print(f"tile {t:2d} @ {r[1] - base:7d}: " + " ".join(f"{names[i]}={r[i] - r[1]}" for i in order[1:]))
