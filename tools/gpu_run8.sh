#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -2 gpurun_out/bench_n1.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/prof_search_1M_r01 -f python tools/quick_search_bench.py --Q 10000 --N 1000000 --iters 1 --check 0 > gpurun_out/ncu_search.log 2>&1
tail -3 gpurun_out/ncu_search.log
