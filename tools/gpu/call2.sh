#!/bin/bash
# GPU call 2: new symmetric kernel (near window, producer-staged thresholds) -- parity tests, then the 1M bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric or golden_neighbour or one_million or row_blocks" > gpurun_out/call2_tests_sym.log 2>&1
echo "sym tests rc=$?"; tail -5 gpurun_out/call2_tests_sym.log
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/call2_tests_multi.log 2>&1
echo "multi tests rc=$?"; tail -5 gpurun_out/call2_tests_multi.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/call2_bench_m1.json 2> gpurun_out/call2_bench_m1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/call2_bench_m1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/call2_bench_m1.json'))
    print(d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['roofline']['frac'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/call2_tests_all.log 2>&1
echo "all tests rc=$?"; tail -5 gpurun_out/call2_tests_all.log
