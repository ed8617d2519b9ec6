"""Where does a small-batch search step go?  Host enqueue time vs device time, per kernel (run under
`ncu --metrics gpu__time_duration.sum` for the per-launch list).

    python tools/small_batch_diag.py [N] [D] [nq] [steps]
"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from two_tower_model_v2_b200 import _native  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    k = 100
    lib = _native.load()
    index, _, _ = bench.make_shard(n_total, d, 1, 0)
    qs = torch.randn((steps + 5, nq, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4321))
    for i in range(5):
        index.search_device(qs[i], k)
    torch.cuda.synchronize()
    # (1) host enqueue time only (the device drains in the background)
    lib.tt_profile_scan_arm(steps)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    t = time.perf_counter()
    for i in range(5, 5 + steps):
        index.search_device(qs[i], k)
    host = (time.perf_counter() - t) / steps * 1e6
    e1.record()
    torch.cuda.synchronize()
    dev = e0.elapsed_time(e1) / steps * 1e3
    sm = torch.empty(steps, dtype=torch.float32)
    n = lib.tt_profile_scan_read(sm.data_ptr(), steps)
    scan = float(sm[:n].mean()) * 1e3
    # (2) one step at a time (device idle between steps): latency of a single search
    lat = []
    for i in range(5, 5 + steps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = index.search_device(qs[i], k)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t) * 1e6)
    lat.sort()
    print(json.dumps({"catalog": f"{n_total}x{d}", "nq": nq, "host_enqueue_us_per_step": host, "device_us_per_step_back_to_back": dev,
                      "main_scan_us": scan, "single_step_latency_us_p50": lat[len(lat) // 2],
                      "launches_per_step": (lib.tt_kernel_launch_count()) / (2 * steps + 5)}))


if __name__ == "__main__":
    main()
