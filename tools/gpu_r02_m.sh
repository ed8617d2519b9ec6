timeout 120 python tools/attn_trace.py 2>&1 | tail -4 | cut -c1-900
timeout 200 python -m pytest -q -p no:cacheprovider --timeout=150 tests/test_gpu_pool.py -x 2>&1 | tail -3
for M in 0 8; do echo -n "mode $M: "; TT_B200_ATTN_MODE=$M timeout 100 python tools/attn_diag.py 2>&1 | tail -1; done
