#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
if [ "$2" != "benchonly" ]; then timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -6; fi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
F='OMP_NUM\|^W\|^\*\*\*\|NCCL version'
timeout 300 $TR --master-port 29701 tools/check_sharded.py 3000000 64 4096 100 2>&1 | grep -v "$F" | tail -4
timeout 600 $TR --master-port 29705 bench.py --gpus $N --steps 40 --warmup 3 > $O/final_n$N.log 2>$O/final_n$N.err
python - <<PY
import json
try:
    r=json.loads([l for l in open(f"$O/final_n$N.log") if l.startswith("{")][-1])
    print(f"N=$N ms/step {r['ms_per_step']:.3f} qps {r['value']:.0f} e2e {r['e2e']['value']:.0f} ({r['e2e']['ms_per_step']:.3f} ms) scan_ms {r['roofline']['kernel_ms']:.3f} unc {r['uncertified_queries']} launches {r['gpu_launches']}", r['config']['sharding'], r['e2e']['api'], r['clocks'])
except Exception as e:
    print("FAILED", e); print(open(f"$O/final_n$N.err").read()[-2500:])
PY
