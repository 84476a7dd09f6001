#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -3 gpurun_out/bench_n1.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -2 gpurun_out/bench_ref.log
nproc; free -g | head -2
