#!/bin/bash
# GPU call 9: CTA-pair symmetric kernel: parity (with a short timeout: a barrier bug would hang), bench, ncu.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric_scan_equals_oracle" > gpurun_out/call9_tests_first.log 2>&1
rc=$?; echo "first sym tests rc=$rc"; tail -5 gpurun_out/call9_tests_first.log
if [ $rc -ne 0 ]; then echo "stopping: pair kernel not correct"; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "symmetric or golden_neighbour or one_million or row_blocks or ties or properties or config" > gpurun_out/call9_tests_sym.log 2>&1
echo "sym tests rc=$?"; tail -3 gpurun_out/call9_tests_sym.log
for opt in "" "--option sym_cta_pair=1"; do
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $opt > gpurun_out/call9_bench_m1.json 2> gpurun_out/call9_bench_m1.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/call9_bench_m1.json'))
    print('m1 [$opt]', d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
done
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/call9_bench_c4.json 2> gpurun_out/call9_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/call9_bench_c4.json')); print('c4', d['ms_per_step'], d['roofline']['executed_frac'])"
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call9_bench_c2.json 2> gpurun_out/call9_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call9_bench_c2.json')); print('c2', d['ms_per_step'], d['stage_ms'], d['config']['scan_symmetric'])"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2e_scan_sym_far_m1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call9_ncu_full.log 2>&1
tail -2 gpurun_out/call9_ncu_full.log
timeout 900 python tools/e2e_host.py --workload m1 --repeat 2 > gpurun_out/r2_e2e_host_m1.json 2> gpurun_out/call9_e2e_host.err
echo "e2e_host m1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_e2e_host_m1.json')); print(d['runs'])"
