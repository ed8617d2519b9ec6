#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider --timeout=400 tests/test_gpu_search.py -x 2>&1 | tail -3
for NQ in 1 128; do
  timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ > $O/r02_small_$NQ.log 2>&1; tail -1 $O/r02_small_$NQ.log
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 24 -c 5 --csv --log-file $O/r02_small_launches_$NQ.csv python tools/small_batch_diag.py 1000000 384 $NQ 12 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_small_launches_$NQ.csv")) if len(r)>5 and r[0].isdigit()]
print("nq=$NQ", [ (r[4].split('(')[0][-30:], round(float(r[-1])/1000,2)) for r in rows[:5]])
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary > $O/r02t_bench.log 2>/dev/null; python -c "
import sys,json
for l in open('gpurun_out/r02t_bench.log'):
    if l.startswith('{'):
        r=json.loads(l); print('headline value', round(r['value']), 'e2e', round(r['e2e']['value']), 'ms', round(r['ms_per_step'],3), 'scan', round(r['roofline']['kernel_ms'],3), 'parity', r['parity']['ok_all_ranks'], 'unc', r['uncertified_queries'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:flat_finalize -c 6 --csv --log-file $O/r02t_fin.csv python bench.py --steps 3 --warmup 3 --no-secondary > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02t_fin.csv")) if len(r)>5 and r[0].isdigit()]
print("finalize nq=4096 us:", [round(float(r[-1])/1000,1) for r in rows])
PY
