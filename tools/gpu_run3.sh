#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core.py -m gpu -q --timeout 300 > gpurun_out/pytest_core.log 2>&1
tail -6 gpurun_out/pytest_core.log
timeout 600 python tools/quick_search_bench.py --Q 10000 --N 100000 > gpurun_out/qb_small.log 2>&1
tail -3 gpurun_out/qb_small.log
timeout 900 python tools/quick_search_bench.py > gpurun_out/qb_full.log 2>&1
tail -3 gpurun_out/qb_full.log
