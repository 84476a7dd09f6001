#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_regions.py > gpurun_out/bench_regions.log 2>&1; tail -4 gpurun_out/bench_regions.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_regions.csv python tools/bench_regions.py --iters 1 --warmup 1 > /dev/null 2>&1
