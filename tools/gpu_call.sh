#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_core.py -m gpu -x -q -k "split_operands" 2>&1 | grep -E "^E|assert|passed|failed" | head -20
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_core.py::test_gemm_nt_split_operands 2>&1 | tail -3
