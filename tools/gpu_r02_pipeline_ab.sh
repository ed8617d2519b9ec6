#!/bin/bash
# pipelined vs serial sharded search at N GPUs (headline workload), N = $1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
for PIPE in 0 1; do
TT_B200_SHARD_PIPELINE=$PIPE timeout 300 $TR --master-port $((29560 + RANDOM % 300)) bench.py --gpus $N --steps 60 --warmup 5 2>/dev/null | tee gpurun_out/r02s_bench_n${N}_pipe$PIPE.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('N=$N pipeline $PIPE: value', round(r['value']), 'ms', round(r['ms_per_step'],3), 'host_enqueue_ms', round(r['host_enqueue_ms_per_step'],3), 'e2e', round(r['e2e']['value']), 'scan_ms', round(r['roofline']['kernel_ms'],3), 'share', round(r['roofline']['kernel_share_of_step'],3), r['clocks'])"
done
