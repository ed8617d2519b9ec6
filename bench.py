#!/usr/bin/env python
"""Benchmark of the serving hot path: exact top-100 inner-product retrieval over a 10M x 384 catalog
(BASELINE.json metric), sharded over N GPUs, plus buyer-tower pooling as a secondary line.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU port of the reference path (rank 0)

A "step" is one search of a batch of `--nq` queries.  Prints ONE JSON line (rank 0).
Catalog (7.68 GB of bf16 per pass, 15.4 GB fp32) is far larger than L2, so no flush is needed between
iterations.  Inputs are synthetic (seeded N(0,1)), generated on the device per shard.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def make_shard(n_total: int, d: int, world: int, rank: int):
    """Catalog rows [lo,hi): i.i.d. N(0,1), seeded per 1M-row block so the catalog is the same for any N."""
    import two_tower_model_v2_b200 as pkg
    lo, hi = pkg.shard_bounds(n_total, world, rank)
    xn = torch.empty((hi - lo, d), device="cuda", dtype=torch.float32)
    blk = 1 << 20
    g = torch.Generator(device="cuda")
    b0 = lo // blk
    r = lo
    while r < hi:
        b = r // blk
        g.manual_seed(1234 + b)
        block = torch.randn((min(blk, n_total - b * blk), d), device="cuda", generator=g)
        s, e = r - b * blk, min(hi, (b + 1) * blk) - b * blk
        xn[r - lo:r - lo + (e - s)].copy_(block[s:e])
        r += e - s
        del block
    idx = pkg.FlatIPIndex.adopt(xn)
    idx.id_offset = lo
    return idx, lo, hi


def bench_pooling(peaks, iters=20):
    """Secondary metric: buyer encodes/s at BASELINE C2 (4096 buyers x 50 events x 384), both modes."""
    import two_tower_model_v2_b200 as pkg
    B, S, D = 4096, 50, 384
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn((B, S, D), device="cuda", generator=g)
    w = torch.tensor([1.0, 5.0, 10.0], device="cuda")[torch.multinomial(torch.tensor([0.75, 0.18, 0.07], device="cuda"), B * S, True, generator=g)].view(B, S)
    out = {}
    alg_bytes = B * S * D * 4 + B * S * 4 + B * D * 4
    for method in ("weighted_avg", "attention"):
        torch.manual_seed(0)
        m = pkg.BuyerTower(D, method).cuda()
        with torch.no_grad():
            for _ in range(3):
                m(x, w)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        e0.record()
        with torch.no_grad():
            for _ in range(iters):
                m(x, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rec = {"buyer_encodes_per_s": B / (ms * 1e-3), "ms_per_batch": ms,
               "hbm_gbs": alg_bytes / (ms * 1e-3) / 1e9, "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        if method == "attention":
            flop = B * S * (2 * D * 128 + 2 * 128)
            # algorithmic flops (2*D*H + 2*H per row) per second; the hidden layer runs as 3xTF32 on the tensor cores
            rec["algorithmic_tflops"] = flop / (ms * 1e-3) / 1e12
            rec["mlp_kernel"] = "tcgen05 kind::tf32, 3xTF32 split (fp32-accurate)"
        out[method] = rec
    out["config"] = "4096 buyers x 50 events x 384 f32 (x 322 MB > L2)"
    return out


def batch_sweep(index, lib, peaks, n_total, d, k, nqs, steps=10):
    """Same search at other query-batch sizes: nq <= ~250 is the HBM-bound regime (catalog streamed once
    per batch), larger batches are tensor-bound.  Device-resident timing, scan kernel timed by events."""
    from two_tower_model_v2_b200 import _native
    dp = int(lib.tt_flat_pitch(d))
    out = []
    for nq in nqs:
        g = torch.Generator(device="cuda").manual_seed(4321 + nq)
        qs = torch.randn((steps + 3, nq, d), device="cuda", generator=g)
        for i in range(3):
            index.search_device(qs[i], k)
        _native.check(lib.tt_profile_scan_arm(steps), "arm")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        nunc = [index.search_device(qs[i], k)[3] for i in range(3, 3 + steps)]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        sm = torch.empty(steps, dtype=torch.float32)
        n = lib.tt_profile_scan_read(sm.data_ptr(), steps)
        scan = float(sm[:n].mean())
        gbs = n_total * dp * 2 / scan / 1e6
        tfl = 2.0 * nq * n_total * d / scan / 1e9
        out.append({"nq": nq, "queries_per_s": nq / ms * 1e3, "ms_per_step": ms, "scan_ms": scan,
                    "scan_hbm_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"], "scan_bf16_tflops": tfl,
                    "tensor_frac_sustained": tfl / peaks["bf16_tflops_sustained"],
                    "uncertified": int(torch.stack(nunc).sum().item())})
    return out


def bench_retrieve_path(index, d, k, requests=200, batch=1024):
    """Secondary (BASELINE config C5 shape on one GPU): the /retrieve path = history rows -> fused gather+pool ->
    exact top-K, host to host.  Batch-1 latency percentiles and batch-1024 throughput, 50-event histories."""
    import two_tower_model_v2_b200 as pkg
    db = pkg.VectorDatabase(d)
    db.index, db.product_ids, db.id_to_index, db.index_to_id = index, [], {}, {}
    out = {}
    for method in ("weighted_avg", "attention"):
        torch.manual_seed(0)
        tower = pkg.BuyerTower(d, method).cuda()
        pipe = pkg.RetrievalPipeline(tower, db)
        S, n = 50, index.ntotal
        g = torch.Generator().manual_seed(99)
        w_choices = torch.tensor([1.0, 5.0, 10.0])

        def make(B):
            idx = torch.randint(0, n, (B, S), generator=g, dtype=torch.int64).pin_memory()
            w = w_choices[torch.multinomial(torch.tensor([0.75, 0.18, 0.07]), B * S, True, generator=g)].view(B, S).pin_memory()
            return idx, w

        def request(idx, w):
            s, i, _ = pipe.retrieve_device_async(idx.cuda(non_blocking=True), w.cuda(non_blocking=True), k).result()
            return s.cpu(), i.cpu()
        reqs = [make(1) for _ in range(requests + 10)]
        for r in reqs[:10]:
            request(*r)
        lat = []
        for r in reqs[10:]:
            t = time.perf_counter()
            request(*r)
            lat.append((time.perf_counter() - t) * 1e3)
        lat = np.sort(np.array(lat))
        big = [make(batch) for _ in range(8)]
        request(*big[0]); request(*big[1])
        torch.cuda.synchronize()
        t = time.perf_counter()
        for r in big[2:]:
            request(*r)
        thr = batch * 6 / (time.perf_counter() - t)
        out[method] = {"batch1_latency_ms": {"p50": float(lat[len(lat) // 2]), "p99": float(lat[int(len(lat) * 0.99) - 1]),
                                             "mean": float(lat.mean())},
                       f"batch{batch}_requests_per_s": thr}
        if method == "weighted_avg":
            # 64 single-request clients (threads) sharing catalog passes through the micro-batcher
            import threading

            def batch_fn(payloads, kk):
                idx = torch.from_numpy(np.stack([p[0] for p in payloads])).pin_memory().cuda(non_blocking=True)
                w = torch.from_numpy(np.stack([p[1] for p in payloads])).pin_memory().cuda(non_blocking=True)
                s, i, _ = pipe.retrieve_device_async(idx, w, kk).result()
                s, i = s.cpu().numpy(), i.cpu().numpy()
                return [list(zip(i[r].tolist(), s[r].tolist())) for r in range(len(payloads))]
            clients, per_client = 64, 30
            payloads = [[(reqs[(c * per_client + j) % len(reqs)][0][0].numpy(), reqs[(c * per_client + j) % len(reqs)][1][0].numpy())
                         for j in range(per_client)] for c in range(clients)]
            lat_mb = []
            with pkg.MicroBatcher(batch_fn, max_batch=128, max_wait_ms=0.3) as mb:
                mb(payloads[0][0], k)

                def client(c):
                    for pl in payloads[c]:
                        t0 = time.perf_counter()
                        mb(pl, k)
                        lat_mb.append((time.perf_counter() - t0) * 1e3)
                ths = [threading.Thread(target=client, args=(c,)) for c in range(clients)]
                t = time.perf_counter()
                [th.start() for th in ths]
                [th.join() for th in ths]
                el = time.perf_counter() - t
                lat_mb = np.sort(np.array(lat_mb))
                out["micro_batched_64_clients"] = {"requests_per_s": clients * per_client / el,
                                                   "mean_batch": (mb.requests - 1) / max(mb.batches - 1, 1),
                                                   "latency_ms": {"p50": float(lat_mb[len(lat_mb) // 2]),
                                                                  "p99": float(lat_mb[int(len(lat_mb) * 0.99) - 1])}}
    out["config"] = f"history 50 events, top-{k}, {index.ntotal}x{d} catalog, host round trip per request, 1 GPU"
    return out


def cpu_baseline_search(n_total, d, nq, k, sample_rows=1 << 20, reps=2):
    """CPU port of the reference search path (faiss-cpu IndexFlatIP is not installable: blocked fp32
    sgemm + top-k, what faiss does for nq >= 20) on a bounded sample of the catalog, all host threads."""
    from oracle import flat_ip_oracle as fo
    sample_rows = min(sample_rows, n_total)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn((sample_rows, d), generator=g)
    x = x / (x.norm(dim=1, keepdim=True) + 1e-8)
    q = torch.randn((nq, d), generator=torch.Generator().manual_seed(4321))
    q = q / (q.norm(dim=1, keepdim=True) + 1e-8)
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fo.torch_search(x, q, k)
        best = min(best, time.perf_counter() - t)
    scale = n_total / sample_rows
    return {"value": nq / (best * scale), "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nq} queries x top-{k} over {sample_rows} of {n_total} rows x {d} f32 (torch sgemm+topk port of "
                      f"IndexFlatIP; faiss-cpu unavailable), time scaled x{scale:g}",
            "seconds_per_sample_step": best, "host_cpus": os.cpu_count()}


def cpu_baseline_pooling(buyers=1024, S=50, D=384, H=128, reps=3):
    """Reported baseline for the second half of the metric (buyer encodes/s): the torch-CPU port of the reference
    BuyerTower (oracle/buyer_tower_oracle.torch_forward: the same eager ops the reference issues) on a bounded sample of
    C2, all host threads."""
    from oracle import buyer_tower_oracle as bo
    g = torch.Generator().manual_seed(99)
    x = torch.randn((buyers, S, D), generator=g)
    w = torch.tensor([1.0, 5.0, 10.0])[torch.multinomial(torch.tensor([0.75, 0.18, 0.07]), buyers * S, True, generator=g)].view(buyers, S)
    torch.manual_seed(0)
    l1, l2 = torch.nn.Linear(D, H), torch.nn.Linear(H, 1)
    params = (l1.weight.detach(), l1.bias.detach(), l2.weight.detach(), l2.bias.detach())
    out = {"cores": torch.get_num_threads(), "kind": "port",
           "sample": f"{buyers} of 4096 buyers x {S} events x {D} f32, torch-CPU eager ops of the reference module"}
    with torch.no_grad():
        for method in ("weighted_avg", "attention"):
            bo.torch_forward(x, w, method, params)
            best = float("inf")
            for _ in range(reps):
                t = time.perf_counter()
                bo.torch_forward(x, w, method, params)
                best = min(best, time.perf_counter() - t)
            out[method] = {"buyer_encodes_per_s": buyers / best, "seconds_per_sample": best}
    return out


def run_reference(args):
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import flat_ip_oracle as fo
    n_total, d, nq, k = args.catalog_rows, args.dim, args.nq, args.topk
    sample_rows = min(1 << 20, n_total)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn((sample_rows, d), generator=g)
    x = x / (x.norm(dim=1, keepdim=True) + 1e-8)
    qs = torch.randn((args.warmup + args.steps, nq, d), generator=torch.Generator().manual_seed(4321))
    for i in range(args.warmup):
        fo.torch_search(x, qs[i] / (qs[i].norm(dim=1, keepdim=True) + 1e-8), k)
    t = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        fo.torch_search(x, qs[i] / (qs[i].norm(dim=1, keepdim=True) + 1e-8), k)
    el = time.perf_counter() - t
    scale = n_total / sample_rows
    ms_step = el / args.steps * 1e3 * scale
    value = nq / (ms_step * 1e-3)
    sample = (f"each step: {nq} queries x top-{k} over {sample_rows} of {n_total} rows (torch-CPU sgemm+topk port of "
              f"faiss IndexFlatIP, faiss-cpu not installable), time scaled x{scale:g}")
    line = {"impl": "reference", "metric": "exact top-100 queries/s, 10Mx384 catalog", "value": value,
            "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"flat_ip_top{k}_{n_total}x{d}_nq{nq}", "catalog_rows": n_total, "dim": d,
                       "topk": k, "query_batch": nq},
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nq", type=int, default=int(os.environ.get("TT_BENCH_NQ", "4096")), help="queries per step")
    ap.add_argument("--catalog-rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--no-secondary", action="store_true", help="skip pooling / cpu baseline extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world, rank, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    import two_tower_model_v2_b200 as pkg
    from two_tower_model_v2_b200 import _native
    lib = _native.load()
    peaks = load_peaks()
    n_total, d, nq, k = args.catalog_rows, args.dim, args.nq, args.topk

    index, lo, hi = make_shard(n_total, d, world, rank)
    n_local = hi - lo
    sharded = pkg.ShardedFlatIPIndex(index, n_total) if world > 1 else None
    total = args.warmup + args.steps
    gq = torch.Generator(device="cuda").manual_seed(4321)
    queries = torch.randn((total, nq, d), device="cuda", generator=gq)
    queries_host = queries.cpu().numpy()

    searcher = sharded if world > 1 else index

    def run_pipelined(submit, first, last):
        """Steps first..last-1, two in flight: step i+1 is enqueued before step i's certificate is looked at, so
        the device never idles on the host.  Every step is checked (uncertified queries re-run through the exact
        path) inside the timed region."""
        bad, pending = 0, None
        for i in range(first, last):
            h = submit(i)
            if pending is not None:
                bad += pending.result()[2]
            pending = h
        if pending is not None:
            bad += pending.result()[2]
        return bad

    # ---- device-resident timing ---------------------------------------------------------------
    run_pipelined(lambda i: searcher.search_async(queries[i], k), 0, args.warmup)
    launches0 = lib.tt_kernel_launch_count()
    _native.check(lib.tt_profile_scan_arm(args.steps), "tt_profile_scan_arm")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    with ClockSampler(local) as clk:
        e0.record()
        uncertified = run_pipelined(lambda i: searcher.search_async(queries[i], k), args.warmup, total)
        e1.record()
        barrier(world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world)
    launches = lib.tt_kernel_launch_count() - launches0
    scan_ms = (torch.empty(args.steps, dtype=torch.float32))
    n_rec = lib.tt_profile_scan_read(scan_ms.data_ptr(), args.steps)
    ms_step = ms_total / args.steps
    value = nq / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (main scan) --------------------------------------------
    dp = int(lib.tt_flat_pitch(d))
    scan_avg_ms = float(scan_ms[:n_rec].mean()) if n_rec > 0 else None
    alg_bytes = n_local * dp * 2 + nq * dp * 2           # bf16 catalog streamed once + the query block(s)
    alg_flops = 2.0 * nq * n_local * d
    crossover = (2 * d / (peaks["hbm_gbs"] * 1e9)) / (2 * d / (peaks["bf16_tflops"] * 1e12))
    roofline = None
    if scan_avg_ms:
        gbs = alg_bytes / (scan_avg_ms * 1e-3) / 1e9
        tfl = alg_flops / (scan_avg_ms * 1e-3) / 1e12
        traffic = None
        tp = ROOT / "profiles" / "scan_traffic.json"      # dram bytes/launch from the committed ncu capture
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get(f"nq{nq}")
            except Exception:
                traffic = None
        if nq < crossover:
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": traffic}
        else:
            roofline = {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": tfl / peaks["bf16_tflops_sustained"], "traffic": traffic}
        roofline.update({"kernel": "flat_scan_kernel<main>", "kernel_ms": scan_avg_ms, "kernel_share_of_step": scan_avg_ms / ms_step,
                         "algorithmic_bytes": alg_bytes, "algorithmic_flops": alg_flops, "hbm_gbs": gbs, "bf16_tflops": tfl,
                         "peaks": peaks["source"], "crossover_nq": crossover})

    # ---- end to end through the host-facing API (numpy in, numpy out) ----------------------------
    if world > 1 and nq % world == 0:
        # every rank fronts 1/G of the batch: its share goes host -> device, the shares are all-gathered over
        # NVLink, and each rank reads back the results of its own rows (whole job: nq*D*4 in, nq*k*12 out)
        nl = nq // world
        host_submit = lambda i: sharded.search_host_sliced_async(queries_host[i][rank * nl:(rank + 1) * nl], k)
        e2e_api = "ShardedFlatIPIndex.search_host_sliced_async(np.ndarray share) -> numpy share, 2 batches in flight"
    else:
        host_submit = lambda i: searcher.search_host_async(queries_host[i], k)
        e2e_api = ("FlatIPIndex" if world == 1 else "ShardedFlatIPIndex") + ".search_host_async(np.ndarray) -> numpy, 2 batches in flight"
    run_pipelined(host_submit, 0, min(2, args.warmup))
    barrier(world)
    t0 = time.perf_counter()
    run_pipelined(host_submit, args.warmup, total)
    barrier(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world) / args.steps
    e2e = {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12 + 4,
           "api": e2e_api}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            sharded.close()
            dist.destroy_process_group()
        return

    line = {"metric": "exact top-100 queries/s, 10Mx384 catalog", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16 scan + f32 rescoring",
            "data": "synthetic",
            "config": {"workload": f"flat_ip_top{k}_{n_total}x{d}_nq{nq}", "catalog_rows": n_total, "dim": d, "topk": k,
                       "query_batch": nq, "sharding": (f"catalog rows over {world} GPUs, one catalog-wide threshold, all-gather ({sharded.exchange_used}) + certified merge"
                                    if world > 1 else "none"),
                       "l2": "inputs (7.68 GB bf16 per pass) exceed L2; no flush"},
            "e2e": e2e, "gpu_launches": int(launches), "uncertified_queries": uncertified,
            "roofline": roofline, "clocks": clk.summary()}
    if world == 1 and not args.no_secondary:
        line["secondary"] = {"pooling": bench_pooling(peaks),
                             "query_batch_sweep": batch_sweep(index, lib, peaks, n_total, d, k, [1, 128, 1024])}
        line["secondary"]["retrieve_path"] = bench_retrieve_path(index, d, k)
        line["cpu_baseline"] = cpu_baseline_search(n_total, d, nq, k)
        try:
            line["secondary"]["pooling"]["cpu_baseline"] = cpu_baseline_pooling()
        except Exception as e:                       # a reported extra must never cost the headline line
            line["secondary"]["pooling"]["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        sharded.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
