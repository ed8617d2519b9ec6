#!/bin/bash
mkdir -p gpurun_out
run() {  # mode, B
  TT_DIAG_B=$2 TT_B200_ATTN_MODE=$1 TT_DIAG_LOGITS_ONLY=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:attn_pool_fused -s 5 -c 4 --csv --log-file gpurun_out/r02k_mode_$1_$2.csv python tools/attn_diag.py > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02k_mode_$1_$2.csv")) if len(r)>5 and r[0].isdigit()]
print("mode $1 B $2 kernel us:", [round(float(r[-1])/1000,1) for r in rows])
PY
}
run 1976 4096; run 4024 4096; run 0 4096; run 2048 4096; run 24 4096; run 2072 4096
for M in 1 2049; do echo -n "mode $M: "; TT_B200_ATTN_MODE=$M timeout 100 python tools/attn_diag.py 2>&1 | tail -1; done
timeout 200 python -m pytest -q -p no:cacheprovider --timeout=150 tests/test_gpu_pool.py -x 2>&1 | tail -3
