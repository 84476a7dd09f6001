#!/bin/bash
# One GPU-box visit that re-validates the round: full GPU test suite, smoke, both bench arms,
# the role timeline of the screen kernel.  Outputs land in gpurun_out/.
#   gpurun --timeout 2700 -- 'bash tools/gpu_validate.sh'
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -40 > gpurun_out/gputest.log; tail -25 gpurun_out/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks', 'parity')})
print(d['roofline'])
print(json.dumps(d['secondary'])[:4000])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/screen_timeline.py --N 125000 > gpurun_out/timeline_125k.txt 2>&1; cat gpurun_out/timeline_125k.txt
timeout 300 python tools/screen_timeline.py --N 1000000 > gpurun_out/timeline_1M.txt 2>&1; cat gpurun_out/timeline_1M.txt
