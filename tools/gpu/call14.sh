#!/bin/bash
# GPU call 14 (2 GPUs): bigger log pool (config 3 must stay symmetric), tail split rule, sanity of everything touched.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu -k "symmetric or gpus or driver or processes or row_blocks" > gpurun_out/call14_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/call14_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call14_bench_m1.json 2> gpurun_out/call14_bench_m1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/call14_bench_m1.json'))
    print('m1', d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['clocks'])
except Exception as e: print("no bench json", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --workload c3 --steps 3 --warmup 2 --no-e2e \
    > gpurun_out/call14_bench_c3_n2.json 2> gpurun_out/call14_bench_c3_n2.err
tail -c 300 gpurun_out/call14_bench_c3_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call14_bench_c3_n2.json') if l.startswith('{')][-1])
    print('c3 n2', d['ms_per_step'], d['stage_ms'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline --option exact_cta_pair=1 > gpurun_out/call14_c5_pair.json 2> gpurun_out/call14_c5.err
python -c "
import json; d=json.load(open('gpurun_out/call14_c5_pair.json')); print('c5 pair', d['ms_per_step'], d['roofline']['frac'])"
