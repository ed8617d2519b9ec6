#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q -p no:cacheprovider -o faulthandler_timeout=100 --timeout=200"
echo "== pool tests"; timeout 300 $PT tests/test_gpu_pool.py > $O/r02f_pool.log 2>&1; echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|AssertionError" $O/r02f_pool.log | head -20
echo "== pooling bench"; timeout 120 python tools/pool_only.py > $O/r02f_poolbench.log 2>&1; echo "rc=$?"; tail -c 700 $O/r02f_poolbench.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file $O/r02f_pool_launches.csv python tools/pool_only.py > $O/r02f_ncu_pool.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02f_pool_launches.csv")) if len(r)>5 and r[0].isdigit()]
agg={}
for r in rows:
    agg.setdefault(r[4][:60],[]).append(float(r[-1]))
for k,v in agg.items():
    if 'tt::' in k: print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.2f} us")
PY
echo "== small batch (in situ per kernel, warm cache)"
for NQ in 1 128; do
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 10 --csv --log-file $O/r02f_small_$NQ.csv python tools/small_batch_diag.py 1000000 384 $NQ 12 > $O/r02f_ncu_small.log 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02f_small_$NQ.csv")) if len(r)>5 and r[0].isdigit()]
print("nq=$NQ", [ (r[4].split('(')[0][-28:], round(float(r[-1])/1000,2)) for r in rows[:10]])
PY
done
for NQ in 1 128; do timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ 2>&1 | tail -1; done
timeout 100 python tools/small_batch_diag.py 10000000 384 1 2>&1 | tail -1
echo "== search tests"; timeout 400 $PT tests/test_gpu_search.py -x > $O/r02f_search.log 2>&1; echo "rc=$?"; tail -3 $O/r02f_search.log
