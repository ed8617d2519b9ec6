#!/bin/bash
# final 2-GPU sanity of the tree: multi-GPU pytest cases, sharded parity at a small and a batch-1 shape, short bench
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 600 python -m pytest -q -p no:cacheprovider --timeout=500 tests/test_gpu_full_size.py -k "sharded" tests/test_gpu_search.py -k "sharded or non_current" 2>&1 | tail -3
timeout 200 $TR --master-port 29512 tools/check_sharded.py 1000000 384 1 100 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -2
timeout 200 $TR --master-port 29513 tools/check_sharded.py 1000000 384 300 100 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -2
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('N=2 value', round(r['value']), 'e2e', round(r['e2e']['value']), 'ms', round(r['ms_per_step'],3), 'parity', r['parity']['ok_all_ranks'], 'unc', r['uncertified_queries'], r['clocks'])"
