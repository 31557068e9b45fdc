#!/bin/bash
# GPU call 19: pool sizing without cudaMemGetInfo on warm calls: parity of the symmetric paths, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric or one_million" > gpurun_out/call19_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/call19_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e > gpurun_out/call19_bench_m1.json 2> gpurun_out/call19_bench_m1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/call19_bench_m1.json'))
print('m1', d['ms_per_step'], d['stage_ms'], d['e2e']['ms'], d['roofline']['frac'], d['clocks'])
PY
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e > gpurun_out/call19_bench_c2.json 2> gpurun_out/call19_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call19_bench_c2.json')); print('c2', d['ms_per_step'], d['stage_ms'], d['e2e']['ms'])"
