#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/profile_sharded.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/profile_sharded_n$NG.txt
