#!/bin/bash
# N-GPU records (N = $1): sharded parity at the benched shape, bench headline / c4 / c5, per-phase times
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 300 $TR --master-port 29511 tools/check_sharded.py 10000000 384 4096 100 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -3 | tee $O/r02p_check_sharded_n$N.log
timeout 300 $TR --master-port 29514 tools/check_sharded_retrieve.py 8000000 384 256 50 100 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -3 | tee -a $O/r02p_check_sharded_n$N.log
for W in headline c4 c5; do
  timeout 600 $TR --master-port $((29560 + RANDOM % 300)) bench.py --gpus $N --workload $W --steps 40 --warmup 4 > $O/r02p_bench_n${N}_$W.log 2> $O/r02p_bench_n${N}_$W.err; echo "$W rc=$?"; grep -v "^W\|^\*\*\*\|^$\|OMP_NUM\|NCCL version" $O/r02p_bench_n${N}_$W.err | tail -5
  python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02p_bench_n${N}_$W.log") if l.startswith("{")][-1])
    rf=r["roofline"] or {}
    print("$W", r["metric"], "value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "ms", round(r["ms_per_step"],3), {k:(round(rf[k],4) if isinstance(rf.get(k),float) else rf.get(k)) for k in ("frac","kernel_ms","kernel_share_of_step")}, "parity", r["parity"]["ok_all_ranks"], "unc", r["uncertified_queries"], r.get("latency_batch1_ms"), r["clocks"])
except Exception as e:
    print("$W bench parse failed", e)
PY
done
timeout 200 $TR --master-port 29541 tools/phase_times.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -8 | tee $O/r02p_phase_times_n$N.log
