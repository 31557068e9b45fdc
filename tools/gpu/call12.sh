#!/bin/bash
# GPU call 12: adaptive near window engaged?  Full GPU test suite, bench, launch list, ncu of the far sweep, API-level e2e.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --option debug_flags=8 > gpurun_out/call12_dbg.json 2> gpurun_out/call12_dbg.err
grep "em2 sym" gpurun_out/call12_dbg.err | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_gpu_1gpu.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/r2_tests_gpu_1gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/call12_bench_m1.json 2> gpurun_out/call12_bench_m1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/call12_bench_m1.json'))
    print('m1', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
except Exception as e: print("no bench json", e)
PY
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/call12_bench_c4.json 2> gpurun_out/call12_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/call12_bench_c4.json')); print('c4', d['ms_per_step'], d['roofline']['executed_frac'])"
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call12_bench_c2.json 2> gpurun_out/call12_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call12_bench_c2.json')); print('c2', d['ms_per_step'], d['stage_ms'], d['e2e']['ms'])"
timeout 900 ncu -k 'regex:[a-z0-9]Kernel' --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_m1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call12_m1_under_ncu.json 2> gpurun_out/call12_ncu_launch.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2f_scan_sym_far_m1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call12_ncu_full.log 2>&1
tail -2 gpurun_out/call12_ncu_full.log
timeout 900 python tools/e2e_host.py --workload m1 --repeat 2 --out gpurun_out/r2_e2e_host_m1.json > gpurun_out/call12_e2e_host.log 2>&1
echo "e2e_host m1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_e2e_host_m1.json')); print(d['runs'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/call12_ref_m1.json 2> gpurun_out/call12_ref_m1.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/call12_ref_m1.json
