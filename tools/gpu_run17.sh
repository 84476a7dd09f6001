#!/bin/bash
# N-GPU bench (whatever the box has) + stage profile
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n$NG.log 2> gpurun_out/bench_n$NG.err; tail -1 gpurun_out/bench_n$NG.log | cut -c1-2500; tail -3 gpurun_out/bench_n$NG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/profile_sharded.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/profile_sharded_n$NG.txt
