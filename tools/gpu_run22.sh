#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_regions.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -4
ISB_TC_DEBUG=8 timeout 300 python tools/tc_ablate.py 8 2>&1 | grep ISB_TC
timeout 600 python tools/bench_regions.py 2>&1 | tail -2 | cut -c1-420 | tee gpurun_out/bench_regions_v3.log
