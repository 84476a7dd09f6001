#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks')})
print(d['roofline'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_pair' --launch-skip 1 -c 1 -f -o gpurun_out/r01_screen_pair_ws python tools/quick_search_bench.py --iters 2 --check 0 > gpurun_out/ncu_pair.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_search_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
