set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -120 > gpurun_out/gputest.log; tail -30 gpurun_out/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 300 python tools/gemm_precision.py > gpurun_out/gemm_precision.json 2>&1; cat gpurun_out/gemm_precision.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks', 'parity')})
print(d['roofline'])
print(json.dumps(d['secondary'])[:6000])
PY
timeout 300 python tools/bench_regions.py --sizes 14,32 2>&1 | tail -2
timeout 300 python tools/bench_mining.py --terms 1 2>&1 | tail -1
for o in "" "screen_wavesync=0" "screen_seed=1" "screen_wavesync=0,screen_seed=1"; do timeout 200 python tools/quick_search_bench.py --N 125000 --check 0 --iters 5 --options "$o" 2>&1 | tail -1; done
