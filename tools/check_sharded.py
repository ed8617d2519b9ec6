#!/usr/bin/env python
"""Multi-GPU parity check of the sharded search (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded.py [N] [D] [nq] [K]

Every rank owns a row shard of the same seeded catalog; the all-gathered + merged top-K must equal, bit for
bit (ids and scores), the top-K a single unsharded index returns on rank 0's GPU for the same queries
(both paths are exact fp32 (score desc, id asc), so equality is the bar, not a tolerance).
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import two_tower_model_v2_b200 as pkg  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    world, rank, local = bench.dist_setup(0)
    torch.cuda.set_device(local)
    index, lo, hi = bench.make_shard(n_total, d, world, rank)
    sharded = pkg.ShardedFlatIPIndex(index, n_total)
    q = torch.randn((nq, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4321))
    s, i, n_bad = sharded.search_device(q, k)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        full, _, _ = bench.make_shard(n_total, d, 1, 0)
        fs, fi, fbad = full.search_checked_device(q, k)
        es, ei = full.search_exact_device(q[:16].contiguous(), k)
        same_ids = bool((fi == i).all())
        same_scores = bool((fs == s).all())
        # exact path: a different fp32 summation order of the query norm -> tie-tolerant comparison (north-star rule)
        exact_ok = bench.compare_topk_device(s[:16], i[:16], es, ei)["ok"]
        nt = min(nq, 256)
        ts, ti = bench.torch_flat_topk(full.xn, q[:nt].contiguous(), k)
        torch_rep = bench.compare_topk_device(s[:nt], i[:nt], ts, ti)
        ok = same_ids and same_scores and exact_ok and torch_rep["ok"]
        print(f"world={world} N={n_total} D={d} nq={nq} K={k}: sharded==single ids {same_ids} scores {same_scores}; "
              f"vs fp32 exact path (16 queries) {exact_ok}; vs torch sgemm+topk ({nt} queries) {torch_rep['ok']} "
              f"(max score err {torch_rep['score_max_abs_err']:.2e}, {torch_rep['queries_with_id_differences']} queries differ inside ties); "
              f"uncertified local {n_bad} single {fbad}; exchange {sharded.exchange_used}"
              + (f" (p2p unavailable: {sharded._p2p_error})" if hasattr(sharded, "_p2p_error") else ""), flush=True)
        del full
    if nq % world == 0:          # sliced host API: every rank submits its share and gets its share back
        nl = nq // world
        hs, hi, _ = sharded.search_host_sliced_async(q.cpu().numpy()[rank * nl:(rank + 1) * nl], k).result()
        same = bool((torch.from_numpy(hi).cuda() == i[rank * nl:(rank + 1) * nl]).all()) and \
            bool((torch.from_numpy(hs).cuda() == s[rank * nl:(rank + 1) * nl]).all())
        ok = ok and same
        if rank == 0:
            print(f"sliced host API == replicated device API on rank 0's share: {same}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
