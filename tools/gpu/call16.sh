#!/bin/bash
# GPU call 16: end-of-round sanity on a fresh box, in the driver's order: smoke, GPU tests, reference arm, bench.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/call16_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/call16_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_gpu_1gpu.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/r2_tests_gpu_1gpu.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/call16_ref.json 2> gpurun_out/call16_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/call16_ref.json
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/call16_bench_m1.json 2> gpurun_out/call16_bench_m1.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/call16_bench_m1.json'))
print('m1', d['ms_per_step'], d['value'], d['stage_ms'], d['e2e']['ms'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])
PY
