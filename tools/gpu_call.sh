#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_pair' --launch-skip 1 -c 1 -f -o gpurun_out/r01_screen_pair_ws python tools/quick_search_bench.py --iters 2 --check 0 > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_search_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'region_pool_fast|region_gather' --launch-skip 6 -c 3 -f -o gpurun_out/r01_regions_v3 python tools/bench_regions.py --sizes 14 --iters 2 --warmup 2 > gpurun_out/ncu_regions.log 2>&1
tail -2 gpurun_out/ncu_regions.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regions.csv python tools/bench_regions.py --sizes 14,32 --iters 2 --warmup 1 > /dev/null 2>&1
timeout 300 python tools/bench_regions.py --sizes 14,32 > gpurun_out/bench_regions_stages.json 2>/dev/null
timeout 300 python tools/bench_mining.py > gpurun_out/bench_mining.json 2>/dev/null
cat gpurun_out/bench_regions_stages.json gpurun_out/bench_mining.json | cut -c1-600
