#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/bench_exchange_stages.py --R 8 2>&1 | tee gpurun_out/exchange_stages_r8.txt
