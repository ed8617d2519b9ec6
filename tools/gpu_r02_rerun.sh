#!/bin/bash
# headline bench at N GPUs after the rank line-up fix of the timed region: the driver's settings and a longer run
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
for ST in 20 60; do
timeout 300 $TR --master-port $((29560 + RANDOM % 300)) bench.py --gpus $N --steps $ST --warmup 5 2>/dev/null | tee gpurun_out/r02w_bench_n${N}_steps$ST.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); rf=r['roofline']; print('N=$N steps $ST: value', round(r['value']), 'ms', round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value']), 'scan_ms', round(rf['kernel_ms'],3), 'share', round(rf['kernel_share_of_step'],3), 'parity', r['parity']['ok_all_ranks'], r['clocks'])"
done
