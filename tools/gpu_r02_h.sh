#!/bin/bash
mkdir -p gpurun_out
for M in 0 1 2 3; do echo -n "mode $M: "; TT_B200_ATTN_MODE=$M timeout 100 python tools/attn_diag.py 2>&1 | tail -1; done
for NQ in 1 128; do timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ 2>&1 | tail -1; done
timeout 200 python -m pytest -q -p no:cacheprovider --timeout=150 tests/test_gpu_pool.py tests/test_gpu_search.py -x 2>&1 | tail -3
