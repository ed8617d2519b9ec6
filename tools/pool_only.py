"""Runs only the buyer-tower pooling benchmark at BASELINE C2 (for an ncu capture of the pooling kernels)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402

print(json.dumps(bench.bench_pooling(bench.load_peaks(), iters=5)))
