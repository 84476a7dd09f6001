#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_mining.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
for seed in 1 0; do
echo "== seed=$seed N=1M"
ISB_SCREEN_SEED=$seed timeout 300 python tools/quick_search_bench.py --iters 8 --check 16 2>&1 | tail -3
echo "== seed=$seed N=125k"
ISB_SCREEN_SEED=$seed timeout 300 python tools/quick_search_bench.py --N 125000 --iters 8 --check 16 2>&1 | tail -3
ISB_SCREEN_SEED=$seed timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:'gemm_tc_pair|rerank' --launch-skip 2 -c 2 python tools/quick_search_bench.py --N 125000 --iters 2 --check 0 2>&1 | grep -E "gemm_tc|rerank_k|gpu__time|cycles_elapsed|tensor_cycles|inst_executed"
done
