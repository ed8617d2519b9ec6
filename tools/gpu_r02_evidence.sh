#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, smoke, bench (+ reference arm), ncu launch list, ncu --set full of the hot
# kernels, query-batch sweeps.  Every ncu pass runs only after the same command has exited 0 without ncu.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02_gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -x -m gpu -p no:cacheprovider 2>&1 | tail -6 > $O/r02_t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 900 python bench.py > $O/r02_bench_n1.log 2> $O/r02_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-secondary"
timeout 600 $B > $O/r02_plain_nq4096.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches_nq4096.csv $B > $O/r02_ncu_launches.log 2>&1
for NQ in 4096 1; do
  C="$B --nq $NQ"
  timeout 600 $C > $O/r02_plain_nq$NQ.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:flat_scan_kernel -s 6 -c 2 -f -o $O/r02_scan_nq$NQ $C > $O/r02_ncu_nq$NQ.log 2>&1
done
timeout 300 python tools/pool_only.py > $O/r02_pool_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_vec_kernel|attn_pool_fused" -s 8 -c 3 -f -o $O/r02_pool_c2 python tools/pool_only.py > $O/r02_ncu_pool.log 2>&1
for NQ in 1 128; do
  timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ > $O/r02_small_$NQ.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 5 --csv --log-file $O/r02_small_launches_$NQ.csv python tools/small_batch_diag.py 1000000 384 $NQ 12 > /dev/null 2>&1
done
timeout 600 python tools/sweep.py 1000000 384 1,128,1024,4096 > $O/r02_sweep_1M.log 2>&1; cp $O/sweep_1000000x384.json $O/r02_sweep_1Mx384.json
timeout 600 python tools/sweep.py 10000000 384 1,64,128,256,512,1024,2048,4096 > $O/r02_sweep_10M.log 2>&1; cp $O/sweep_10000000x384.json $O/r02_sweep_10Mx384.json
timeout 600 python tools/sweep.py 10000000 768 1,64,128,256,1024,4096 > $O/r02_sweep_10M_768.log 2>&1; cp $O/sweep_10000000x768.json $O/r02_sweep_10Mx768.json
tail -n 3 $O/r02_t_gpu.log $O/r02_smoke.log; tail -c 600 $O/r02_bench_n1.log; echo; tail -c 400 $O/r02_bench_reference.log; echo; cat $O/r02_small_1.log $O/r02_small_128.log; ls -la $O | grep r02_ | head -40
