#!/bin/bash
# GPU call 5: threshold warp + parallel pivot acceptance: parity, bench, launch list, ncu of the far sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric or golden_neighbour or one_million or row_blocks or ties or properties" > gpurun_out/call5_tests_sym.log 2>&1
echo "sym tests rc=$?"; tail -3 gpurun_out/call5_tests_sym.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/call5_bench_m1.json 2> gpurun_out/call5_bench_m1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/call5_bench_m1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/call5_bench_m1.json'))
    print('m1', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['roofline']['frac'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call5_bench_c2.json 2> gpurun_out/call5_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call5_bench_c2.json')); print('c2', d['ms_per_step'], d['stage_ms'])"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/call5_bench_c4.json 2> gpurun_out/call5_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/call5_bench_c4.json')); print('c4', d['ms_per_step'], d['stage_ms'])"
timeout 900 ncu -k 'regex:[a-z0-9]Kernel' --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_bench_m1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call5_m1_under_ncu.json 2> gpurun_out/call5_ncu_launch.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2c_scan_sym_far_m1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call5_ncu_full.log 2>&1
tail -2 gpurun_out/call5_ncu_full.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/call5_tests_all.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/call5_tests_all.log
