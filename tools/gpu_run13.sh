#!/bin/bash
# round-1 re-entry check: GPU tests, smoke, bench (own arm + reference arm), region/mining stage benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.log | cut -c1-1500; tail -3 gpurun_out/bench_n1.err
timeout 600 python tools/bench_regions.py > gpurun_out/bench_regions.log 2>&1; tail -2 gpurun_out/bench_regions.log
