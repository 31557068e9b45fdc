#!/bin/bash
# GPU call 3: leader grouping + near window: parity, then where the 1M job's candidates come from.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric or golden_neighbour or one_million or row_blocks or ties or properties" > gpurun_out/call3_tests_sym.log 2>&1
echo "sym tests rc=$?"; tail -5 gpurun_out/call3_tests_sym.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --option debug_flags=8 > gpurun_out/call3_dbg.json 2> gpurun_out/call3_dbg.err
grep "em2 sym" gpurun_out/call3_dbg.err | tail -4
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/call3_bench_m1.json 2> gpurun_out/call3_bench_m1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/call3_bench_m1.err
python - <<'PY'
import json
for f in ('gpurun_out/call3_bench_m1.json',):
    try:
        d=json.load(open(f))
        print(d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['roofline']['frac'], d['config']['scan_symmetric'], d['e2e'] and d['e2e']['stage_ms'])
    except Exception as e: print("no bench json", e)
PY
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/call3_bench_c4.json 2> gpurun_out/call3_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/call3_bench_c4.json')); print('c4', d['ms_per_step'], d['stage_ms'])"
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call3_bench_c2.json 2> gpurun_out/call3_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call3_bench_c2.json')); print('c2', d['ms_per_step'], d['stage_ms'])"
timeout 600 python bench.py --workload c2 --symmetric --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call3_bench_c2sym.json 2> gpurun_out/call3_bench_c2sym.err
python -c "
import json; d=json.load(open('gpurun_out/call3_bench_c2sym.json')); print('c2 sym', d['ms_per_step'], d['stage_ms'])"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/call3_tests_all.log 2>&1
echo "all tests rc=$?"; tail -5 gpurun_out/call3_tests_all.log
