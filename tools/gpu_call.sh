#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_regions.py --sizes 32 2>&1 | tail -12 | cut -c1-400
