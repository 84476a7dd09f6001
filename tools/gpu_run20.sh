#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_regions.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/bench_regions.py --sizes 14 2>&1 | tail -2 | tee gpurun_out/bench_regions_tc.log
