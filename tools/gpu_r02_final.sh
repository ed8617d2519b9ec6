#!/bin/bash
# refresh of the one-GPU records after the last kernel changes: GPU tests, smoke, default bench, small-batch step
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu -p no:cacheprovider 2>&1 | tail -4 > $O/r02_t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 900 python bench.py > $O/r02_bench_n1.log 2> $O/r02_bench_n1.err; echo "bench rc=$?"
for NQ in 1 128; do
  timeout 100 python tools/small_batch_diag.py 1000000 384 $NQ > $O/r02_small_$NQ.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 24 -c 5 --csv --log-file $O/r02_small_launches_$NQ.csv python tools/small_batch_diag.py 1000000 384 $NQ 12 > /dev/null 2>&1
done
timeout 300 python tools/sweep.py 1000000 384 1,128,1024,4096 > $O/r02_sweep_1M.log 2>&1; cp $O/sweep_1000000x384.json $O/r02_sweep_1Mx384.json
tail -n 2 $O/r02_t_gpu.log $O/r02_smoke.log; tail -c 300 $O/r02_bench_n1.log; echo; cat $O/r02_small_1.log $O/r02_small_128.log
