#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_pool.py -q -x 2>&1 | tail -4
timeout 300 python tools/pool_only.py 2>&1 | tail -1
