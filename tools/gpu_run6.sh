#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mining.py tests/test_gpu_core.py -m gpu -q --timeout 300 > gpurun_out/pytest_mining.log 2>&1
tail -40 gpurun_out/pytest_mining.log
