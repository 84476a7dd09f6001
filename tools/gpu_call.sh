#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_regions.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -8
V14="ISB_GATHER_G=1,ISB_GATHER_STAGES=1;ISB_GATHER_G=2,ISB_GATHER_STAGES=2;ISB_POOL_STAGES=3;ISB_POOL_STAGES=4,ISB_GATHER_G=8;ISB_GATHER_G=8,ISB_GATHER_STAGES=4,ISB_GATHER_CB=16;ISB_GATHER_G=2,ISB_GATHER_STAGES=2,ISB_GATHER_CB=16;ISB_GATHER_G=1,ISB_GATHER_STAGES=1,ISB_GATHER_CB=64"
timeout 600 python tools/bench_regions.py --sizes 14 --variants "$V14" > gpurun_out/regions_variants14.jsonl 2> gpurun_out/regions_variants14.err
V32="ISB_GATHER_G=1,ISB_GATHER_STAGES=1;ISB_GATHER_G=1,ISB_GATHER_STAGES=1,ISB_GATHER_CB=8;ISB_GATHER_G=8,ISB_GATHER_CB=8;ISB_GATHER_G=4,ISB_GATHER_CB=8,ISB_GATHER_STAGES=2;ISB_GATHER_G=2,ISB_GATHER_CB=16,ISB_GATHER_STAGES=2"
timeout 600 python tools/bench_regions.py --sizes 32 --variants "$V32" > gpurun_out/regions_variants32.jsonl 2> gpurun_out/regions_variants32.err
cat gpurun_out/regions_variants14.jsonl gpurun_out/regions_variants32.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'][36:44], d['variant'], {k: round(v, 4) for k, v in d['ms'].items()})
"
ISB_GATHER_G=1 ISB_GATHER_STAGES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'region_gather' --launch-skip 4 -c 1 -f -o gpurun_out/r01_gather_g1 python tools/bench_regions.py --sizes 14 --iters 2 --warmup 2 > gpurun_out/ncu_gather.log 2>&1
tail -2 gpurun_out/ncu_gather.log
