#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"region_pool|region_gather" -s 2 -c 2 -o gpurun_out/prof_regions14 -f python tools/bench_regions.py --iters 1 --warmup 1 --sizes 14 > gpurun_out/ncu_regions.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"region_pool|region_gather" -s 2 -c 2 -o gpurun_out/prof_regions32 -f python tools/bench_regions.py --iters 1 --warmup 1 --sizes 32 >> gpurun_out/ncu_regions.log 2>&1
tail -3 gpurun_out/ncu_regions.log
