#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q -p no:cacheprovider -o faulthandler_timeout=100 --timeout=170"
echo "== pool tests"; timeout 500 $PT tests/test_gpu_pool.py --durations=5 > $O/r02c_pool.log 2>&1; echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|assert [0-9]|AssertionError: \(" $O/r02c_pool.log | head -40
echo "== pooling bench"; timeout 120 python tools/pool_only.py > $O/r02c_poolbench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02c_pool_launches.csv python tools/pool_only.py > $O/r02c_ncu_pool.log 2>&1
echo "rc=$?"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02c_pool_launches.csv")) if len(r)>5 and r[0].isdigit()]
agg={}
for r in rows:
    name=r[4][:60]; v=float(r[-1])
    agg.setdefault(name,[]).append(v)
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.2f} us  last={v[-1]/1000:.2f}")
PY
echo "== small batch ncu"; ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file $O/r02c_small_launches.csv python tools/small_batch_diag.py 1000000 384 1 12 > $O/r02c_ncu_small.log 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02c_small_launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[:12]: print(f"{r[4][:70]:70s} grid={r[7]} {float(r[-1])/1000:8.2f} us")
PY
