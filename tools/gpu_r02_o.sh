#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest -q -p no:cacheprovider --timeout=200 tests/test_gpu_pool.py -x 2>&1 | tail -8
timeout 120 python tools/attn_trace.py 2>&1 | sed -n '8,13p' | cut -c1-900
for M in 0 8; do echo -n "mode $M: "; TT_B200_ATTN_MODE=$M timeout 100 python tools/attn_diag.py 2>&1 | tail -1; done
TT_B200_ATTN_MODE=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:attn_pool_fused -s 5 -c 6 --csv --log-file $O/r02o_fused.csv python tools/pool_only.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02o_fused.csv")) if len(r)>5 and r[0].isdigit()]
print("fused kernel us:", [round(float(r[-1])/1000,1) for r in rows])
PY
echo "== search"; timeout 400 python -m pytest -q -p no:cacheprovider --timeout=300 tests/test_gpu_search.py -x 2>&1 | tail -4
for NQ in 1 128; do timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ 2>&1 | tail -1; done
for NQ in 1 128; do
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 5 --csv --log-file $O/r02o_small_$NQ.csv python tools/small_batch_diag.py 1000000 384 $NQ 12 > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02o_small_$NQ.csv")) if len(r)>5 and r[0].isdigit()]
print("nq=$NQ", [ (r[4].split('(')[0][-28:], round(float(r[-1])/1000,2)) for r in rows[:5]])
PY
done
