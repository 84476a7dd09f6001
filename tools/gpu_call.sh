#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_mining.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/quick_search_bench.py --iters 8 --check 16 2>&1 | tail -3
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:'gemm_tc' --launch-skip 1 -c 1 python tools/quick_search_bench.py --iters 2 --check 0 2>&1 | grep -E "dram__bytes|gpu__time|cycles_elapsed|tensor_cycles|inst_executed"
echo "== shard-sized (N=125000)"
timeout 300 python tools/quick_search_bench.py --N 125000 --iters 8 --check 16 2>&1 | tail -3
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:'gemm_tc' --launch-skip 1 -c 1 python tools/quick_search_bench.py --N 125000 --iters 2 --check 0 2>&1 | grep -E "dram__bytes|gpu__time|cycles_elapsed|tensor_cycles|inst_executed"
