#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core.py -m gpu -q --timeout 300 > gpurun_out/pytest_core.log 2>&1
tail -15 gpurun_out/pytest_core.log
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
cat gpurun_out/gemm_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/prof_search_r1a -f python tools/quick_search_bench.py --Q 10000 --N 100000 --iters 1 --check 0 > gpurun_out/ncu_search.log 2>&1
tail -5 gpurun_out/ncu_search.log
ls -la gpurun_out
