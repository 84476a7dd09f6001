#!/bin/bash
# first GPU contact: core parity tests + search probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_core.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_core.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_core.log
tail -40 gpurun_out/pytest_core.log
timeout 600 python tools/quick_search_bench.py --Q 1024 --N 100000 --D 2048 > gpurun_out/qb_small.log 2>&1
tail -12 gpurun_out/qb_small.log
timeout 900 python tools/quick_search_bench.py > gpurun_out/qb_full.log 2>&1
tail -12 gpurun_out/qb_full.log
