#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q -p no:cacheprovider -o faulthandler_timeout=100 --timeout=200"
echo "== pool tests"; timeout 300 $PT tests/test_gpu_pool.py > $O/r02d_pool.log 2>&1; echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|AssertionError" $O/r02d_pool.log | head -20
echo "== search tests (TS scan on)"; timeout 600 $PT tests/test_gpu_search.py tests/test_gpu_full_size.py -x --durations=8 > $O/r02d_search.log 2>&1; echo "rc=$?"; tail -22 $O/r02d_search.log
echo "== bench TS on/off"
for TS in 1 0; do TT_B200_SCAN_TS=$TS timeout 200 python bench.py --no-secondary --steps 10 --warmup 3 > $O/r02d_bench_ts$TS.log 2> $O/r02d_bench_ts$TS.err; echo "TS=$TS rc=$?"; tail -c 300 $O/r02d_bench_ts$TS.err; python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02d_bench_ts$TS.log") if l.startswith("{")][-1])
    print("value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "ms", round(r["ms_per_step"],3), {k:round(r["roofline"][k],4) for k in ("frac_sustained","frac_burst","kernel_ms","kernel_share_of_step")}, "parity", r["parity"]["ok_all_ranks"], "unc", r["uncertified_queries"], r["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
done
echo "== sweep TS on/off"
for TS in 1 0; do echo TS=$TS; TT_B200_SCAN_TS=$TS timeout 200 python tools/sweep.py 10000000 384 300,512,1024,2048 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['nq'], 'ms', round(r['ms_per_step'],3), 'scan', round(r['scan_ms'],3), 'tensor_frac_sus', round(r['tensor_frac_sustained'],3), 'unc', r['uncertified'])"; done
TT_B200_SCAN_TS=1 timeout 200 python tools/sweep.py 1000000 384 1,128,1024,4096 2>&1 | grep '^{' | cut -c1-200
TT_B200_SCAN_TS=1 timeout 200 python tools/sweep.py 10000000 768 1024,4096 2>&1 | grep '^{' | cut -c1-200
echo "== ncu fused attention"
timeout 120 python tools/pool_only.py > $O/r02d_poolbench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_pool_fused -s 2 -c 1 -o $O/r02d_attn_fused python tools/pool_only.py > $O/r02d_ncu_attn.log 2>&1; echo "rc=$?"; tail -c 600 $O/r02d_poolbench.log
echo "== ncu finalize/select small batch"
ncu --set full --clock-control none --import-source on -k regex:"flat_finalize|select_threshold" -s 20 -c 2 -o $O/r02d_small python tools/small_batch_diag.py 1000000 384 1 12 > $O/r02d_ncu_small.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep | tail -3
