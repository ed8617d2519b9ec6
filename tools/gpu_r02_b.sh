#!/bin/bash
# Round-2 diagnostic run: find the slow/hanging test, first run of the fused attention kernel, small-batch step anatomy.
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q -x -p no:cacheprovider -o faulthandler_timeout=100 --timeout=170"
echo "== duplicates/clustered"; timeout 240 $PT tests/test_gpu_search.py -k "duplicates or clustered or exact_path_alone" -v > $O/r02b_dup.log 2>&1; echo "rc=$?"; tail -25 $O/r02b_dup.log
echo "== pool tests"; timeout 420 $PT tests/test_gpu_pool.py -v --durations=8 > $O/r02b_pool.log 2>&1; echo "rc=$?"; tail -40 $O/r02b_pool.log
echo "== full size C3"; timeout 300 $PT tests/test_gpu_full_size.py -k C3 -s > $O/r02b_c3.log 2>&1; echo "rc=$?"; grep -E "^\[|passed|failed|Error|error" $O/r02b_c3.log | tail -20
echo "== pooling bench"; timeout 120 python tools/pool_only.py > $O/r02b_poolbench.log 2>&1; echo "rc=$?"; tail -c 1500 $O/r02b_poolbench.log
echo "== small batch"; for P in 1 0; do TT_B200_PDL=$P timeout 120 python tools/small_batch_diag.py 1000000 384 1 2>&1 | tail -1; TT_B200_PDL=$P timeout 120 python tools/small_batch_diag.py 1000000 384 128 2>&1 | tail -1; done
TT_B200_PDL=1 timeout 150 python tools/small_batch_diag.py 10000000 384 1 2>&1 | tail -1
echo "== bench"; timeout 240 python bench.py --no-secondary --steps 10 --warmup 3 > $O/r02b_bench.log 2> $O/r02b_bench.err; echo "rc=$?"; tail -c 400 $O/r02b_bench.err; python - <<'PY'
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02b_bench.log") if l.startswith("{")][-1])
    print("value", r["value"], "e2e", r["e2e"]["value"], "ms", r["ms_per_step"], {k:r["roofline"][k] for k in ("frac","frac_sustained","frac_burst","kernel_ms","kernel_share_of_step")}, "parity", r["parity"]["ok_all_ranks"], r["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
