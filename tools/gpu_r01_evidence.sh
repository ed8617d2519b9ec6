#!/bin/bash
# Round-1 evidence run on one B200: GPU tests, smoke, bench, ncu launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -15 > $O/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_default.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-secondary"
timeout 600 $B > $O/plain_nq4096.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_nq4096.csv $B > $O/ncu_launches.log 2>&1
for NQ in 4096 128 1; do
  C="$B --nq $NQ"
  timeout 600 $C > $O/plain_nq$NQ.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:flat_scan_kernel -s 6 -c 2 -f -o $O/scan_nq$NQ $C > $O/ncu_nq$NQ.log 2>&1
done
C="$B --nq 4096"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flat_finalize_kernel|select_threshold_kernel" -s 6 -c 2 -f -o $O/finalize_nq4096 $C > $O/ncu_finalize.log 2>&1
timeout 300 python tools/pool_only.py > $O/pool_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_vec_kernel|attn_logits" -s 8 -c 3 -f -o $O/pool_c2 python tools/pool_only.py > $O/ncu_pool.log 2>&1
timeout 600 python tools/sweep.py 1000000 384 > $O/sweep_1M.log 2>&1
timeout 600 python tools/sweep.py 10000000 384 > $O/sweep_10M.log 2>&1
timeout 600 python tools/sweep.py 10000000 768 1,64,128,1024,1024,4096 > $O/sweep_10M_768.log 2>&1
tail -n 3 $O/t_gpu.log $O/smoke.log; tail -c 1500 $O/bench_default.log; tail -n 2 $O/bench_reference.log; ls -la $O
