#!/usr/bin/env python
"""Per-phase device times of the sharded search (CUDA events on each rank, max over ranks printed by rank 0):

    torchrun --nproc-per-node G tools/phase_times.py [N] [D] [nq] [K] [steps]
"""
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import two_tower_model_v2_b200 as pkg  # noqa: E402
from two_tower_model_v2_b200 import _native  # noqa: E402
from two_tower_model_v2_b200.sharded import record_layout, record_views  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    world, rank, local = bench.dist_setup(0)
    torch.cuda.set_device(local)
    index, lo, hi = bench.make_shard(n_total, d, world, rank)
    sh = pkg.ShardedFlatIPIndex(index, n_total)
    qs = torch.randn((steps + 3, nq, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4321))
    lay, rec, gathered = sh._buffers(nq, k, qs.device, world)
    scores, ids, bound, flags = record_views(rec, lay, nq, k)
    topr = torch.empty((nq, _native.TT_SHARD_TOPR), device="cuda")
    topr_g = torch.empty((world, nq, _native.TT_SHARD_TOPR), device="cuda")
    p2p = sh._p2p_for(nq, k, qs.device, world, lay)
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    names = ["sample", "allgather_topr", "search(select+scan+finalize)", "allgather_records", "merge"]
    acc = torch.zeros(len(names), dtype=torch.float64)
    for it in range(steps + 3):
        q = qs[it]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        index.shard_sample(q, k, n_total, topr)
        ev[1].record()
        if p2p is not None:
            topr_g = p2p.all_gather(0, topr.view(torch.uint8).view(-1), status).view(torch.float32).view(world, nq, _native.TT_SHARD_TOPR)
        else:
            dist.all_gather_into_tensor(topr_g.view(-1), topr.view(-1))
        ev[2].record()
        index.shard_search_into(nq, k, n_total, topr_g, scores, ids, bound, flags)
        ev[3].record()
        if p2p is not None:
            gathered = p2p.all_gather(1, rec, status)
        else:
            dist.all_gather_into_tensor(gathered.view(-1), rec)
        ev[4].record()
        sh._merge(gathered, lay, nq, k)
        ev[5].record()
        torch.cuda.synchronize()
        if it >= 3:
            acc += torch.tensor([ev[j].elapsed_time(ev[j + 1]) for j in range(len(names))], dtype=torch.float64)
    acc /= steps
    t = acc.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} N={n_total} D={d} nq={nq} K={k} record={lay.nbytes/1e6:.2f} MB/rank exchange={sh.exchange_used} status={status.tolist()}")
        for n, v in zip(names, t.tolist()):
            print(f"  {n:32s} {v*1e3:9.1f} us (max over ranks, barrier before each step)")
        print(f"  {'sum':32s} {sum(t.tolist())*1e3:9.1f} us")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
