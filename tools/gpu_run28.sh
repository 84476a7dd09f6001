#!/bin/bash
# N-GPU: contract bench, stage profile, NCCL test, config-5 end to end
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n$NG.log 2> gpurun_out/bench_n$NG.err; tail -1 gpurun_out/bench_n$NG.log | cut -c1-300; tail -2 gpurun_out/bench_n$NG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/profile_sharded.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/profile_sharded_n$NG.txt
timeout 600 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/e2e_instance_search.py > gpurun_out/e2e_n$NG.log 2> gpurun_out/e2e_n$NG.err; tail -1 gpurun_out/e2e_n$NG.log; tail -2 gpurun_out/e2e_n$NG.err
