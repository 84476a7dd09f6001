#!/bin/bash
# full GPU check: all parity tests, smoke, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -2 gpurun_out/bench_n1.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|rerank|f32_to_bf16|fill_u32|merge" -c 400 --csv --log-file gpurun_out/launches_search.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/prof_search_r01 -f python tools/quick_search_bench.py --Q 10000 --N 200000 --iters 1 --check 0 > gpurun_out/ncu_search.log 2>&1
tail -3 gpurun_out/ncu_search.log
ls -la gpurun_out
