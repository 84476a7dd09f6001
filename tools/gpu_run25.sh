#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/e2e_instance_search.py > gpurun_out/e2e_n$NG.log 2> gpurun_out/e2e_n$NG.err; tail -1 gpurun_out/e2e_n$NG.log; tail -4 gpurun_out/e2e_n$NG.err
