#!/bin/bash
# round-1 final single-GPU verification: tests, smoke, both bench arms, launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.log | cut -c1-1200; tail -3 gpurun_out/bench_n1.err
timeout 600 python tools/bench_regions.py > gpurun_out/bench_regions.log 2>&1; tail -2 gpurun_out/bench_regions.log | cut -c1-400
timeout 600 python tools/bench_mining.py > gpurun_out/bench_mining.log 2>&1; tail -1 gpurun_out/bench_mining.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_search_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regions.csv python tools/bench_regions.py --iters 2 --warmup 1 > gpurun_out/ncu_regions.log 2>&1
