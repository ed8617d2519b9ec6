#!/bin/bash
# fused attention v3: parity tests, timing (events + per-kernel ncu)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest -q -p no:cacheprovider --timeout=200 tests/test_gpu_pool.py -x 2>&1 | tail -15
for M in 0 4 8; do echo -n "mode $M: "; TT_B200_ATTN_MODE=$M timeout 100 python tools/attn_diag.py 2>&1 | tail -1; done
timeout 120 python tools/pool_only.py 2>&1 | tail -c 600; echo
for M in 0 4 8; do
TT_B200_ATTN_MODE=$M timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:attn_pool_fused -s 5 -c 6 --csv --log-file $O/r02l_mode_$M.csv python tools/pool_only.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02l_mode_$M.csv")) if len(r)>5 and r[0].isdigit()]
print("mode $M fused kernel us:", [round(float(r[-1])/1000,1) for r in rows])
PY
done
