#!/bin/bash
# GPU call 6: elected-lane MMA issue: parity, bench, ncu of the far sweep; API-level e2e (c2); config 5 with recall.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric or golden_neighbour or one_million or row_blocks or ties or properties" > gpurun_out/call6_tests_sym.log 2>&1
echo "sym tests rc=$?"; tail -3 gpurun_out/call6_tests_sym.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/call6_bench_m1.json 2> gpurun_out/call6_bench_m1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/call6_bench_m1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/call6_bench_m1.json'))
    print('m1', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['roofline']['frac'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
for L in 1024 512; do
timeout 600 python bench.py --workload c4 --lsh $L --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/call6_bench_c4_$L.json 2> gpurun_out/call6_bench_c4_$L.err
python -c "
import json; d=json.load(open('gpurun_out/call6_bench_c4_$L.json')); print('c4 L=$L', d['ms_per_step'], d['roofline']['executed_frac'])"
done
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call6_bench_c2.json 2> gpurun_out/call6_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call6_bench_c2.json')); print('c2', d['ms_per_step'], d['stage_ms'])"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2d_scan_sym_far_m1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call6_ncu_full.log 2>&1
tail -2 gpurun_out/call6_ncu_full.log
timeout 600 python tools/e2e_host.py --workload c2 > gpurun_out/r2_e2e_host_c2.json 2> gpurun_out/call6_e2e_host.err
echo "e2e_host rc=$?"; cat gpurun_out/r2_e2e_host_c2.json | cut -c1-900; tail -c 300 gpurun_out/call6_e2e_host.err
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r2_bench_c5.json 2> gpurun_out/call6_c5.err
echo "c5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c5.json')); print('c5', d['ms_per_step'], d['roofline']['frac'], d.get('recall'))"
