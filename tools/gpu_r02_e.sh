#!/bin/bash
# 2-GPU development run: pipelined sharded search + sharded /retrieve parity, bench at N=2 (headline, c4, c5), fused attention v2.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
echo "== fused attention v2"; timeout 200 python -m pytest -q -p no:cacheprovider --timeout=150 tests/test_gpu_pool.py 2>&1 | tail -4
timeout 100 python tools/pool_only.py 2>&1 | tail -c 700
echo "== check_sharded"; timeout 300 $TR --master-port 29511 tools/check_sharded.py 10000000 384 4096 100 2>&1 | grep -v "^W\|^\*\*\*" | tail -4
timeout 200 $TR --master-port 29512 tools/check_sharded.py 1000000 384 300 100 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
timeout 200 $TR --master-port 29513 tools/check_sharded.py 3000 64 50 1000 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
echo "== check_sharded_retrieve"; timeout 300 $TR --master-port 29514 tools/check_sharded_retrieve.py 4000000 384 256 50 100 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
echo "== bench N=2"
for W in headline c4 c5; do
  timeout 400 $TR --master-port 2952$RANDOM bench.py --gpus 2 --workload $W --steps 12 --warmup 4 > $O/r02e_bench_n2_$W.log 2> $O/r02e_bench_n2_$W.err; echo "$W rc=$?"; grep -v "^W\|^\*\*\*\|^$" $O/r02e_bench_n2_$W.err | tail -5
  python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02e_bench_n2_$W.log") if l.startswith("{")][-1])
    rf=r["roofline"] or {}
    print("$W", r["metric"], "value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "ms", round(r["ms_per_step"],3), {k:(round(rf[k],4) if isinstance(rf.get(k),float) else rf.get(k)) for k in ("frac","kernel_ms","kernel_share_of_step")}, "parity", r["parity"]["ok_all_ranks"], "unc", r["uncertified_queries"], r.get("latency_batch1_ms"))
except Exception as e:
    print("$W bench parse failed", e)
PY
done
TT_B200_SHARD_PIPELINE=0 timeout 300 $TR --master-port 29531 bench.py --gpus 2 --steps 12 --warmup 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('pipeline off: value', round(r['value']), 'e2e', round(r['e2e']['value']), 'share', round(r['roofline']['kernel_share_of_step'],3))"
