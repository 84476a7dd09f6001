#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o gpurun_out/prof_search_1M_r01b -f python tools/quick_search_bench.py --Q 10000 --N 1000000 --D 2048 --k 100 > gpurun_out/ncu_search.log 2>&1
tail -2 gpurun_out/ncu_search.log | cut -c1-300
