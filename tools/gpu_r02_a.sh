#!/bin/bash
# Round-2 first evidence run on one B200: GPU tests (incl. full-size parity), default bench, reference arm.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt 2>&1
free -g >> $O/gpu.txt; nproc >> $O/gpu.txt
timeout 1500 python -m pytest tests -q -x -m gpu --durations=15 > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 $O/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02a_bench.log 2> $O/r02a_bench.err; echo "bench rc=$?"
tail -c 1500 $O/r02a_bench.err
python - <<'PY'
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02a_bench.log") if l.startswith("{")][-1])
    print("value", r["value"], "e2e", r["e2e"]["value"], "ms", r["ms_per_step"], "roof", {k:r["roofline"][k] for k in ("frac","frac_sustained","frac_burst","kernel_ms","kernel_share_of_step")})
    print("parity", r["parity"]["ok_all_ranks"], r["parity"]["vs_exact_fp32_path"], r["parity"]["vs_torch_fp32_sgemm_topk"])
    print("clocks", r["clocks"])
    s=r["secondary"]
    print("pooling", {m:(s["pooling"][m]["ms_per_batch"], s["pooling"][m]["hbm_frac"]) for m in ("weighted_avg","attention")})
    for e in s["query_batch_sweep"]+s.get("query_batch_sweep_1Mx384",[]): print(e["catalog"], e["nq"], "ms", round(e["ms_per_step"],4), "scan", round(e["scan_ms"],4), "frac_of_floor", round(e["step_frac_of_floor"],3))
    print("retrieve", s["retrieve_path"])
    print("cpu", r["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02a_bench.log").read()[-2000:])
PY
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02a_ref.log 2>&1; echo "ref rc=$?"; tail -c 1200 $O/r02a_ref.log
