#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -q 2>&1 | tail -60 > gpurun_out/t_search.log
timeout 900 python tools/sweep.py 10000000 384 > gpurun_out/sweep.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/ncu_bench.log 2>&1
tail -n 5 gpurun_out/t_search.log; tail -n 12 gpurun_out/sweep.log; tail -n 2 gpurun_out/bench_full.log
