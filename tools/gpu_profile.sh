#!/bin/bash
# One GPU-box visit that collects the round's ncu evidence (one GPU; never a multi-rank command):
# launch lists (gpu__time_duration) of the bench and the region head, and --set full captures of the
# dominant kernels.  Reports land in gpurun_out/ and are summarised here with tools/ncu_summary.py.
#   gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'
set -x
mkdir -p gpurun_out
R=r02
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/${R}_launches_search_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -c 260 --csv --log-file gpurun_out/${R}_launches_regions.csv \
    python tools/bench_regions.py --sizes 14 --iters 2 --warmup 1 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -c 140 --csv --log-file gpurun_out/${R}_launches_regions32.csv \
    python tools/bench_regions.py --sizes 32 --iters 2 --warmup 1 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/${R}_launches_mining.csv \
    python tools/bench_mining.py --iters 2 > /dev/null 2>&1
FULL="$NCU --set full --import-source on -f"
timeout 600 $FULL -k regex:gemm_tc_pair_kernel --launch-skip 1 -c 1 -o gpurun_out/${R}_screen_1M \
    python tools/quick_search_bench.py --iters 1 --check 0 > gpurun_out/ncu_screen_1M.log 2>&1
timeout 600 $FULL -k regex:gemm_tc_pair_kernel --launch-skip 1 -c 1 -o gpurun_out/${R}_screen_125k \
    python tools/quick_search_bench.py --N 125000 --iters 1 --check 0 > gpurun_out/ncu_screen_125k.log 2>&1
timeout 600 $FULL -k regex:'region_pool_fast|region_gather|region_candidates|region_finalize|region_logits' --launch-skip 9 -c 6 \
    -o gpurun_out/${R}_regions_14 python tools/bench_regions.py --sizes 14 --iters 2 --warmup 2 > gpurun_out/ncu_regions14.log 2>&1
timeout 600 $FULL -k regex:'region_pool_fast|region_gather|region_candidates' --launch-skip 8 -c 4 \
    -o gpurun_out/${R}_regions_32 python tools/bench_regions.py --sizes 32 --iters 2 --warmup 2 > gpurun_out/ncu_regions32.log 2>&1
timeout 600 $FULL -k regex:'gemm_tc_pair_kernel|pool_candidates_kernel|mining_rerank' --launch-skip 3 -c 3 \
    -o gpurun_out/${R}_mining python tools/bench_mining.py --iters 2 > gpurun_out/ncu_mining.log 2>&1
ls -la gpurun_out/*.ncu-rep
