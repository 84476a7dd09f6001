#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regions.py -m gpu -q --timeout 300 > gpurun_out/pytest_regions.log 2>&1
tail -30 gpurun_out/pytest_regions.log
