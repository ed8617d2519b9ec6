"""GPU: parity AT THE BENCHED SIZES (BASELINE.json metric line and configs C3 / C4), through the C-ABI.

For each catalog (generated on the device exactly as bench.py does) and query batch:
  * every query is certified (or re-run) and its list is sorted;
  * ALL queries against an independent fp32 checker on the device (cuBLAS sgemm with TF32 off + torch.topk,
    bench.torch_flat_topk) under the north-star tolerance: scores within 1e-5, ids identical except inside
    score ties of 1e-6;
  * >= 64 sampled queries against this library's always-exact fp32 path (tt_flat_search_exact);
  * >= 8 sampled queries against the CPU oracle (oracle/flat_ip_oracle.py) fed by chunked device->host copies
    of the stored rows, ties classified on fp64 scores.
Reference semantics: IndexFlatIP.search as called at src/inference/vector_db.py:160,197.
"""
import numpy as np
import pytest
import torch

import bench
from oracle import flat_ip_oracle as fo

pytestmark = pytest.mark.gpu

K = 100
CASES = [
    ("metric_10Mx384", 10_000_000, 384, [4096]),
    ("C3_1Mx384", 1_000_000, 384, [4096, 1]),
    ("C4_10Mx768", 10_000_000, 768, [64, 128, 4096]),
]


class _DeviceRows:
    """xn[rows] for compare_topk's fp64 tie classification without a host copy of the whole table."""

    def __init__(self, xn):
        self.xn = xn

    def __getitem__(self, rows):
        return self.xn[torch.as_tensor(np.asarray(rows, dtype=np.int64), device=self.xn.device)].cpu().numpy()


def oracle_topk_chunked(xn_dev, qn, k, chunk=1 << 20):
    """fo.search over device-resident rows, 1M-row chunks through host memory; (score desc, id asc) merge."""
    cs, ci = [], []
    for lo in range(0, xn_dev.shape[0], chunk):
        xc = xn_dev[lo:lo + chunk].cpu().numpy()
        s, i = fo.search(xc, qn, k, block=1 << 18)
        cs.append(s)
        ci.append(i + lo)
    cs, ci = np.concatenate(cs, 1), np.concatenate(ci, 1)
    out_s, out_i = np.empty((qn.shape[0], k), np.float32), np.empty((qn.shape[0], k), np.int64)
    for r in range(qn.shape[0]):
        order = np.lexsort((ci[r], -cs[r]))[:k]
        out_s[r], out_i[r] = cs[r][order], ci[r][order]
    return out_s, out_i


@pytest.mark.parametrize("name,N,D,nqs", CASES, ids=[c[0] for c in CASES])
def test_benched_sizes_match_exact_path_torch_checker_and_oracle(name, N, D, nqs):
    import time
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    index, _, _ = bench.make_shard(N, D, 1, 0)
    dev = index.xn.device
    torch.cuda.synchronize()

    def lap(what):
        nonlocal t0
        torch.cuda.synchronize()
        t = time.perf_counter()
        print(f"[{name}] {what}: {t - t0:.2f} s", flush=True)
        t0 = t
    lap("catalog")
    try:
        for nq in nqs:
            q = torch.randn((nq, D), device=dev, generator=torch.Generator(device="cuda").manual_seed(4321))
            s, i, flags, nunc = index.search_device(q, K)
            n_bad = int(nunc.item())
            if n_bad:
                index.search_exact_device(q, K, s, i, torch.nonzero(flags != 1).flatten().to(torch.int32))
            # i.i.d. catalogs: the planner's thresholds must certify (nearly) everything on the tensor-core path
            assert n_bad <= max(1, nq // 1000), f"{name} nq={nq}: {n_bad} queries fell back to the exact path"
            assert bool((s[:, 1:] <= s[:, :-1]).all()) and int(i.min()) >= 0 and int(i.max()) < N
            lap(f"nq={nq} search")
            # (1) every query vs the independent fp32 checker
            ts, ti = bench.torch_flat_topk(index.xn, q, K)
            rep = bench.compare_topk_device(s, i, ts, ti)
            assert rep["ok"], f"{name} nq={nq} vs torch sgemm+topk: {rep}"
            lap(f"nq={nq} torch checker ({rep['queries_with_id_differences']} queries differ inside ties)")
            # (2) sampled queries vs the library's exact fp32 path
            sel = torch.linspace(0, nq - 1, min(nq, 64), device=dev).long()
            es, ei = index.search_exact_device(q[sel].contiguous(), K)
            rep = bench.compare_topk_device(s[sel], i[sel], es, ei)
            assert rep["ok"], f"{name} nq={nq} vs tt_flat_search_exact: {rep}"
            lap(f"nq={nq} exact path")
            # (3) sampled queries vs the CPU oracle
            sub = torch.linspace(0, nq - 1, min(nq, 8), device=dev).long()
            qn = fo.normalize_rows(q[sub].cpu().numpy())
            rs, ri = oracle_topk_chunked(index.xn, qn, K)
            ok, msg = fo.compare_topk(s[sub].cpu().numpy(), i[sub].cpu().numpy(), rs, ri, _DeviceRows(index.xn), qn)
            assert ok, f"{name} nq={nq} vs CPU oracle: {msg}"
            lap(f"nq={nq} CPU oracle")
    finally:
        del index
        torch.cuda.empty_cache()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("N,D,nq", [(10_000_000, 384, 4096), (10_000_000, 768, 1024)])
def test_sharded_on_real_gpus_equals_single_index(N, D, nq):
    """Hardware multi-GPU parity (skipped on a 1-GPU box): tools/check_sharded.py under torchrun, one rank per
    GPU - the all-gathered + merged result must equal the single-index result bit for bit and the exact path."""
    import subprocess
    import sys
    from pathlib import Path
    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip("needs >= 2 GPUs")
    G = 8 if G >= 8 else (4 if G >= 4 else 2)
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(G),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        str(root / "tools" / "check_sharded.py"), str(N), str(D), str(nq), str(K)],
                       capture_output=True, text=True, cwd=root, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "sharded==single ids True scores True" in r.stdout and "vs fp32 exact path (16 queries) True" in r.stdout


def test_sharded_retrieve_path_on_real_gpus_equals_single_device():
    """BASELINE config C5 shape on real GPUs (skipped on a 1-GPU box): owner-computes pooling over the row-sharded item
    table + sharded exact top-K == the single-device /retrieve pipeline (tools/check_sharded_retrieve.py)."""
    import subprocess
    import sys
    from pathlib import Path
    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip("needs >= 2 GPUs")
    G = 8 if G >= 8 else (4 if G >= 4 else 2)
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(G),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        str(root / "tools" / "check_sharded_retrieve.py"), "4000000", "384", "256", "50", str(K)],
                       capture_output=True, text=True, cwd=root, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert r.stdout.count("ok=True") >= 4
