#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q -p no:cacheprovider -o faulthandler_timeout=100 --timeout=200"
echo "== pool tests"; timeout 300 $PT tests/test_gpu_pool.py > $O/r02g_pool.log 2>&1; echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|AssertionError" $O/r02g_pool.log | head -20
echo "== pooling bench"; timeout 120 python tools/pool_only.py > $O/r02g_poolbench.log 2>&1; echo "rc=$?"; tail -c 700 $O/r02g_poolbench.log; echo
ncu --set full --clock-control none --import-source on -k regex:attn_pool_fused -s 2 -c 1 -o $O/r02g_attn_fused python tools/pool_only.py > $O/r02g_ncu_attn.log 2>&1; echo "ncu rc=$?"
echo "== small batch (in situ per kernel, warm cache)"
for NQ in 1 128; do
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 5 --csv --log-file $O/r02g_small_$NQ.csv python tools/small_batch_diag.py 1000000 384 $NQ 12 > $O/r02g_ncu_small.log 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02g_small_$NQ.csv")) if len(r)>5 and r[0].isdigit()]
print("nq=$NQ", [ (r[4].split('(')[0][-28:], round(float(r[-1])/1000,2)) for r in rows[:5]])
PY
done
for NQ in 1 128; do timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ 2>&1 | tail -1; done
timeout 100 python tools/small_batch_diag.py 10000000 384 1 2>&1 | tail -1
timeout 200 python tools/sweep.py 10000000 768 64,128,256 2>&1 | grep '^{' | cut -c1-260
echo "== search tests"; timeout 400 $PT tests/test_gpu_search.py -x > $O/r02g_search.log 2>&1; echo "rc=$?"; tail -3 $O/r02g_search.log
