#!/bin/bash
# P2P exchange bring-up on N GPUs: parity check, phase times (p2p vs nccl), short bench.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
F='OMP_NUM\|^W\|^\*\*\*\|NCCL version'
timeout 300 $TR --master-port 29601 tools/check_sharded.py 1000000 384 300 100 2>&1 | grep -v "$F" | tail -6
timeout 300 $TR --master-port 29602 tools/check_sharded.py 3000000 64 4096 100 2>&1 | grep -v "$F" | tail -3
timeout 300 $TR --master-port 29603 tools/phase_times.py 2>&1 | grep -v "$F" | tail -8
TT_B200_EXCHANGE=nccl timeout 300 $TR --master-port 29604 tools/phase_times.py 2>&1 | grep -v "$F" | tail -8
timeout 600 $TR --master-port 29605 bench.py --gpus $N --steps 40 --warmup 3 > $O/p2p_n$N.log 2>$O/p2p_n$N.err
TT_B200_EXCHANGE=nccl timeout 600 $TR --master-port 29606 bench.py --gpus $N --steps 40 --warmup 3 > $O/nccl_n$N.log 2>$O/nccl_n$N.err
python - <<PY
import json
for tag in ["p2p","nccl"]:
    try:
        r=json.loads([l for l in open(f"$O/{tag}_n$N.log") if l.startswith("{")][-1])
        print(tag, f"N=$N ms/step {r['ms_per_step']:.3f} qps {r['value']:.0f} e2e {r['e2e']['value']:.0f} scan_ms {r['roofline']['kernel_ms']:.3f} unc {r['uncertified_queries']} launches {r['gpu_launches']}", r['config']['sharding'])
    except Exception as e:
        print(tag, "FAILED", e); print(open(f"$O/{tag}_n$N.err").read()[-2000:])
PY
