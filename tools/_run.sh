mkdir -p gpurun_out
: > gpurun_out/per_file.log
rm -f gpurun_out/parity_achieved.jsonl
for f in tests/test_gpu_core.py tests/test_gpu_regions.py tests/test_gpu_generic_windows.py tests/test_gpu_mining.py tests/test_gpu_sharded.py tests/test_gpu_torch_ops.py tests/test_gpu_dropin.py tests/test_gpu_baseline_sizes.py tests/test_gpu_config0_resnet152.py; do
  echo "=== $f" >> gpurun_out/per_file.log
  timeout -k 10 400 python -m pytest $f -m gpu -q --timeout 150 2>&1 | grep -v "mbarrier wait" | tail -60 >> gpurun_out/per_file.log
  echo "=== $f"; tail -4 gpurun_out/per_file.log
done
