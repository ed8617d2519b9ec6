#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest -q -p no:cacheprovider --timeout=200 tests/test_gpu_pool.py -x 2>&1 | tail -12
for M in 0 8 16; do echo -n "mode $M: "; TT_B200_ATTN_MODE=$M timeout 100 python tools/attn_diag.py 2>&1 | tail -1; done
TT_B200_ATTN_MODE=0 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --cache-control none -k regex:attn_pool_pair -s 5 -c 4 --csv --log-file $O/r02q_fused.csv python tools/pool_only.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02q_fused.csv")) if len(r)>5 and r[0].isdigit()]
print("fused kernel:", [(r[-3], r[-2], r[-1]) for r in rows])
PY
