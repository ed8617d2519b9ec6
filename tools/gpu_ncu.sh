#!/bin/bash
# ncu --set full capture of the main scan kernel at nq=$1 (default 128); plain run first.
NQ=${1:-128}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-secondary --nq $NQ"
$CMD > gpurun_out/plain_nq$NQ.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flat_scan_kernel -s 6 -c 4 -f -o gpurun_out/scan_nq$NQ $CMD > gpurun_out/ncu_nq$NQ.log 2>&1
tail -n 3 gpurun_out/ncu_nq$NQ.log
