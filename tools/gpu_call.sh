#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks')})
print(d['roofline'])
print(json.dumps(d['secondary'], indent=1))
"
tail -n 5 gpurun_out/bench_n1.err
