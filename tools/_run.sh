set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -80 > gpurun_out/gputest_first.log; tail -60 gpurun_out/gputest_first.log
for f in tests/test_gpu_*.py; do timeout 600 python -m pytest $f -m gpu -q -x 2>&1 | tail -3; done
