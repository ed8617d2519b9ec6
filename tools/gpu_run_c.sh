#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_search.py -q -x -k "scan_scores or matches_oracle" 2>&1 | tail -6
timeout 600 python tools/sweep.py 10000000 768 128,256,1024,4096 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print('10Mx768 nq', r['nq'], 'ms', round(r['ms_per_step'],3), 'qps', round(r['qps']), 'scan', round(r['scan_ms'],3), 'hbm', round(r['hbm_frac'],2), 'tens', round(r['tensor_frac_sustained'],2), 'unc', r['uncertified'])
"
