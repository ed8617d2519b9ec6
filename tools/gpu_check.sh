#!/bin/bash
# Runs the GPU test suites in separate processes (a trapped kernel poisons its CUDA context) and a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_pool.py -q -x 2>&1 | tail -40 > gpurun_out/t_pool.log
timeout 600 python -m pytest tests/test_gpu_search.py -q -k "scan_scores" 2>&1 | tail -80 > gpurun_out/t_scan.log
timeout 1500 python -m pytest tests/test_gpu_search.py -q -k "not scan_scores" 2>&1 | tail -120 > gpurun_out/t_search.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/t_pool.log gpurun_out/t_scan.log gpurun_out/t_search.log gpurun_out/smoke.log gpurun_out/bench.log
