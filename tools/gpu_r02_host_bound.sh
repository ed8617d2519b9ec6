#!/bin/bash
# is the pipelined sharded search host-bound at short steps?  2 ranks over a 2.5M-row catalog (1.25M rows per rank, as at 8 x 10M)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
for PIPE in 1 0; do
TT_B200_SHARD_PIPELINE=$PIPE timeout 300 $TR --master-port $((29560 + RANDOM % 300)) bench.py --gpus 2 --catalog-rows 2500000 --steps 60 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('pipeline $PIPE: value', round(r['value']), 'ms', round(r['ms_per_step'],3), 'host_enqueue_ms', round(r['host_enqueue_ms_per_step'],3), 'e2e', round(r['e2e']['value']), 'scan_ms', round(r['roofline']['kernel_ms'],3), 'share', round(r['roofline']['kernel_share_of_step'],3), 'launches/step', r['gpu_launches']/r['steps'])"
done
