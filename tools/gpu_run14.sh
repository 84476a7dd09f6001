#!/bin/bash
# sharded search with candidate exchange: GPU tests (1 GPU, lock-step shards) + N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_core.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; tail -15 gpurun_out/pytest_sharded.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n$NG.log 2> gpurun_out/bench_n$NG.err; tail -1 gpurun_out/bench_n$NG.log | cut -c1-2500; tail -3 gpurun_out/bench_n$NG.err
fi
