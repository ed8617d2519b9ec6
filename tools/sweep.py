"""Query-batch sweep of the exact search on one GPU (device-resident timing) -> gpurun_out/sweep.json."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import two_tower_model_v2_b200 as pkg  # noqa: E402
from two_tower_model_v2_b200 import _native  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    nqs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 16, 64, 128, 256, 512, 1024, 2048, 4096]
    k = 100
    lib = _native.load()
    peaks = bench.load_peaks()
    index, lo, hi = bench.make_shard(n_total, d, 1, 0)
    dp = int(lib.tt_flat_pitch(d))
    rows = []
    for nq in nqs:
        steps = 10 if nq <= 1024 else 4
        g = torch.Generator(device="cuda").manual_seed(4321 + nq)
        qs = torch.randn((steps + 3, nq, d), device="cuda", generator=g)
        for i in range(3):
            index.search_device(qs[i], k)
        lib.tt_profile_scan_arm(steps)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        nunc = []
        flags_all = []
        e0.record()
        for i in range(3, 3 + steps):
            out = index.search_device(qs[i], k)
            nunc.append(out[3]); flags_all.append(out[2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        sm = torch.empty(steps, dtype=torch.float32)
        n = lib.tt_profile_scan_read(sm.data_ptr(), steps)
        scan = float(sm[:n].mean())
        rec = {"nq": nq, "ms_per_step": ms, "qps": nq / ms * 1e3, "scan_ms": scan,
               "scan_hbm_gbs": n_total * dp * 2 / scan / 1e6, "scan_tflops": 2.0 * nq * n_total * d / scan / 1e9,
               "hbm_frac": n_total * dp * 2 / scan / 1e6 / peaks["hbm_gbs"],
               "tensor_frac_sustained": 2.0 * nq * n_total * d / scan / 1e9 / peaks["bf16_tflops_sustained"],
               "uncertified": int(torch.stack(nunc).sum().item()),
               "reasons": sorted(set((-f[f != 1]).tolist()) if False else torch.unique(-torch.cat(flags_all)[torch.cat(flags_all) != 1]).tolist())}
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"sweep_{n_total}x{d}.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
