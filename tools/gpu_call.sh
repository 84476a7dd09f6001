#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_regions.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_regions.py --sizes 14,32 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'][36:44], {k: round(v, 4) for k, v in d['ms'].items()})
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_regions.csv python tools/bench_regions.py --sizes 14,32 --iters 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/launches_regions.csv')) if len(r) > 14 and r[0].isdigit()]
seen = {}
for r in rows:
    if 'isb::' in r[4]:
        key = (r[4][:60], r[8], r[7])
        seen.setdefault(key, []).append(int(r[14]))
for k, v in seen.items():
    print(k[0].replace('void ', ''), k[1], k[2], sorted(v)[len(v)//2], len(v))
PY
