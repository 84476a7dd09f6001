set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -m gpu -q --durations=5 -x 2>&1 | tail -150 > gpurun_out/gputest.log; tail -5 gpurun_out/gputest.log
timeout 600 python -m pytest tests/test_gpu_core.py::test_topk_search_clustered_and_planted "tests/test_gpu_mining.py" tests/test_gpu_config0_resnet152.py tests/test_gpu_torch_ops.py -q 2>&1 | tail -150 > gpurun_out/gputest2.log; tail -5 gpurun_out/gputest2.log
timeout 300 python tools/gemm_precision.py > gpurun_out/gemm_precision.json 2>&1; cat gpurun_out/gemm_precision.json
timeout 300 python tools/bench_regions.py --sizes 14,32 2>&1 | tail -3
