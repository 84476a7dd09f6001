#!/bin/bash
# One GPU-box visit that re-validates the round: full GPU test suite, smoke, both bench arms,
# launch lists and ncu captures of the dominant kernels.  Outputs land in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh'
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks')})
print(d['roofline'])
print(json.dumps(d['secondary'])[:3000])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_search_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'region_pool_fast|region_gather' --launch-skip 6 -c 3 -f -o gpurun_out/r01_regions_v4 python tools/bench_regions.py --sizes 14 --iters 2 --warmup 2 > gpurun_out/ncu_regions.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regions.csv python tools/bench_regions.py --sizes 14,32 --iters 2 --warmup 1 > /dev/null 2>&1
timeout 300 python tools/bench_regions.py --sizes 14,32 > gpurun_out/bench_regions_stages.json 2>/dev/null
timeout 300 python tools/bench_mining.py > gpurun_out/bench_mining.json 2>/dev/null
