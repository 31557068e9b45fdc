#!/bin/bash
# GPU call 11 (2 GPUs): adaptive near window -- parity on 1 and 2 GPUs, bench at 1 and 2 GPUs, pair vs single CTAs at c2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_multi.py -x -q -m gpu -k "symmetric or golden_neighbour or one_million or row_blocks or ties or properties or config or gpus or driver or processes" > gpurun_out/call11_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/call11_tests.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --option debug_flags=8 > gpurun_out/call11_dbg.json 2> gpurun_out/call11_dbg.err
grep "em2 sym" gpurun_out/call11_dbg.err | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call11_bench_m1.json 2> gpurun_out/call11_bench_m1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/call11_bench_m1.json'))
    print('m1', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['roofline']['frac'])
except Exception as e: print("no bench json", e)
PY
for opt in "" "--option sym_cta_pair=1" "--one-directional"; do
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $opt > gpurun_out/call11_bench_c2.json 2> gpurun_out/call11_bench_c2.err
python - <<PY
import json
d=json.load(open('gpurun_out/call11_bench_c2.json')); print('c2 [$opt]', d['ms_per_step'], d['stage_ms'], d['config']['scan_symmetric'])
PY
done
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/call11_bench_c4.json 2> gpurun_out/call11_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/call11_bench_c4.json')); print('c4', d['ms_per_step'], d['roofline']['executed_frac'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/call11_bench_m1_n2.json 2> gpurun_out/call11_bench_m1_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call11_bench_m1_n2.json') if l.startswith('{')][-1])
    print('m1 n2', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['e2e'] and d['e2e']['stage_ms'].get('lists_equal_device_path'))
except Exception as e: print("no bench json", e)
PY
