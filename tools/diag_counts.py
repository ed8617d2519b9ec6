"""Distribution of per-query candidate counts / thresholds (diagnostic)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from two_tower_model_v2_b200 import _native, ops

n_total, d, nq, k = int(sys.argv[1]), 384, int(sys.argv[2]), 100
lib = _native.load()
index, lo, hi = bench.make_shard(n_total, d, 1, 0)
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 7
g = torch.Generator(device="cuda").manual_seed(seed)
qs_all = torch.randn((int(sys.argv[3]) if len(sys.argv) > 3 else 8, nq, d), device="cuda", generator=g)
allc, allf = [], []
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 8):
    q = qs_all[it]
    s, i, flags, nunc = index.search_device(q, k)
    ws = index._workspace(nq, k)
    thr = torch.empty(nq, device="cuda"); cnt = torch.empty(nq, device="cuda", dtype=torch.int32)
    _native.check(lib.tt_flat_debug_read(ws.data_ptr(), n_total, d, nq, k, thr.data_ptr(), cnt.data_ptr(), ops._stream()), "dbg")
    allc.append(cnt.clone()); allf.append(flags.clone())
    bad = torch.nonzero(flags != 1).flatten()
    for b in bad.tolist():
        qq = q[b] / q[b].norm(); qb = qq.to(torch.bfloat16).float(); dq = float((qb - qq).norm()); st = index.stats.tolist()
        print("uncertified it", it, "q", b, "flag", int(flags[b]), "count", int(cnt[b]), "thr", float(thr[b]), "kth score", float(s[b, k - 1]), "top", float(s[b, 0]), "dq", dq, "stats", st[:2])
c = torch.cat(allc).float()
print("queries", c.numel(), "count mean", c.mean().item(), "std", c.std().item(), "min", c.min().item(), "max", c.max().item())
print("quantiles", torch.quantile(c, torch.tensor([0.0001, 0.001, 0.01, 0.1, 0.5, 0.9, 0.99, 0.999], device="cuda")).tolist())
print("uncertified", int((torch.cat(allf) != 1).sum()))
