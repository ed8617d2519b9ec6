#!/bin/bash
# Multi-GPU check + scaling bench on one box: usage  tools/gpu_multi.sh "2 4 8"   (GPU counts to run)
NS=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-secondary > $O/scale_n1.log 2>&1
for N in $NS; do
  timeout 600 $TR --nproc-per-node $N --master-port 2951$N tools/check_sharded.py 1000000 384 300 100 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
  timeout 600 $TR --nproc-per-node $N --master-port 2952$N tools/check_sharded.py 3000 64 50 1000 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
  timeout 900 $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 40 --warmup 3 > $O/scale_n$N.log 2>$O/scale_n$N.err
done
python - <<PY
import json
base=None
for n in [1]+[int(x) for x in "$NS".split()]:
    try:
        r=json.loads([l for l in open(f"$O/scale_n{n}.log") if l.startswith("{")][-1])
        if n==1: base=r["value"]
        rf=r["roofline"]
        print(f"N={n} ms/step {r['ms_per_step']:.3f} qps {r['value']:.0f} x{r['value']/base:.2f} e2e {r['e2e']['value']:.0f} scan_ms {rf['kernel_ms']:.3f} share {rf['kernel_share_of_step']:.3f} unc {r['uncertified_queries']} launches {r['gpu_launches']} clocks {r['clocks']}")
    except Exception as e:
        print("N", n, "FAILED", e)
        import subprocess; print(open(f"$O/scale_n{n}.log").read()[-800:]); 
        try: print(open(f"$O/scale_n{n}.err").read()[-1500:])
        except Exception: pass
PY
