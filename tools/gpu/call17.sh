#!/bin/bash
# GPU call 17: the bench line with the API-level leg; host-layer GPU tests after the device-first reordering.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_layer.py -x -q -m gpu > gpurun_out/call17_tests_host.log 2>&1
echo "host tests rc=$?"; tail -2 gpurun_out/call17_tests_host.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/call17_bench_m1.json 2> gpurun_out/call17_bench_m1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/call17_bench_m1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/call17_bench_m1.json'))
print('m1', d['ms_per_step'], d['e2e']['ms'], d.get('e2e_api'))
PY
