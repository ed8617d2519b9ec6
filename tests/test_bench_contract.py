"""bench.py contract checks that need no GPU: the reference arm (CPU port) prints the agreed JSON line, and the
product arm refuses to run without a CUDA device (there is no CPU fallback to time by accident)."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--catalog-rows", "20000", "--nq", "16", "--topk", "10")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["value"] > 0
    assert line["config"]["workload"] == "flat_ip_top10_20000x384_nq16"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0


def test_product_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--catalog-rows", "20000", "--nq", "16")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
