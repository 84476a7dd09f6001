#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"region_pool_tc" -s 2 -c 1 -o gpurun_out/prof_pool_tc -f python tools/bench_regions.py --iters 1 --warmup 2 --sizes 14 > gpurun_out/ncu_pool_tc.log 2>&1
tail -3 gpurun_out/ncu_pool_tc.log
ls -la gpurun_out/*.ncu-rep
