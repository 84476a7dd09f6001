#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"region_gather" -s 4 -c 1 -o gpurun_out/prof_gather -f python tools/bench_regions.py --iters 1 --warmup 2 --sizes 14 > gpurun_out/ncu_gather.log 2>&1
tail -2 gpurun_out/ncu_gather.log | cut -c1-200
