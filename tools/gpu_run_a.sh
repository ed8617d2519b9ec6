#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -6
timeout 600 python tools/sweep.py 10000000 768 1,64,128,256,1024,4096 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print('10Mx768 nq', r['nq'], 'ms', round(r['ms_per_step'],3), 'qps', round(r['qps']), 'scan', round(r['scan_ms'],3), 'hbm', round(r['hbm_frac'],2), 'tens', round(r['tensor_frac_sustained'],2), 'unc', r['uncertified'])
"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_a.log 2>$O/bench_a.err || tail -20 $O/bench_a.err
python - <<PY
import json
r=json.loads([l for l in open("$O/bench_a.log") if l.startswith("{")][-1])
print({k:r[k] for k in ['value','ms_per_step','gpu_launches','uncertified_queries']}, 'e2e', r['e2e']['value'])
print(json.dumps(r['secondary']['retrieve_path'], indent=1))
PY
