#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_mining.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_n1_quick.log 2> gpurun_out/bench_n1_quick.err; tail -1 gpurun_out/bench_n1_quick.log | cut -c1-1800; tail -3 gpurun_out/bench_n1_quick.err
