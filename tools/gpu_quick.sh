#!/bin/bash
# Quick GPU iteration: tests, a short bench at a few batch sizes, per-kernel launch list at nq=4096.
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -15 > $O/t_gpu.log
tail -n 6 $O/t_gpu.log
for NQ in 4096 1024 256 128 1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --nq $NQ > $O/q_nq$NQ.log 2>&1
  python - <<PY
import json
try:
    r=json.loads(open("$O/q_nq$NQ.log").read().strip().splitlines()[-1])
    rf=r["roofline"]
    print("nq", $NQ, "ms/step %.3f"%r["ms_per_step"], "qps %.0f"%r["value"], "e2e %.0f"%r["e2e"]["value"], "scan_ms %.3f"%rf["kernel_ms"], "hbm %.0f"%rf["hbm_gbs"], "tflops %.0f"%rf["bf16_tflops"], "unc", r["uncertified_queries"], r["clocks"])
except Exception as e:
    print("nq", $NQ, "FAILED", e); print(open("$O/q_nq$NQ.log").read()[-1500:])
PY
done
B="python bench.py --steps 3 --warmup 3 --no-secondary"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_nq4096.csv $B > $O/ncu_launches.log 2>&1
python tools/summarise_launches.py $O/launches_nq4096.csv --skip 22 | tail -n 12
