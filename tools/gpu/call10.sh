#!/bin/bash
# GPU call 10 (8 GPUs): the metric's workload and config 3 on 8 GPUs, multi-GPU parity tests on >2 GPUs.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 \
    > gpurun_out/call10_bench_m1_n8.json 2> gpurun_out/call10_bench_m1_n8.err
echo "bench m1 n8 rc=$?"; tail -c 800 gpurun_out/call10_bench_m1_n8.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call10_bench_m1_n8.json') if l.startswith('{')][-1])
    print('m1 n8', d['ms_per_step'], d['stage_ms'], d['e2e'], d['roofline']['frac'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 5 --warmup 3 \
    > gpurun_out/call10_bench_m1_n4.json 2> gpurun_out/call10_bench_m1_n4.err
echo "bench m1 n4 rc=$?"; tail -c 400 gpurun_out/call10_bench_m1_n4.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call10_bench_m1_n4.json') if l.startswith('{')][-1])
    print('m1 n4', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 8 --workload c3 --steps 5 --warmup 3 \
    > gpurun_out/call10_bench_c3_n8.json 2> gpurun_out/call10_bench_c3_n8.err
echo "bench c3 n8 rc=$?"; tail -c 400 gpurun_out/call10_bench_c3_n8.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call10_bench_c3_n8.json') if l.startswith('{')][-1])
    print('c3 n8', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
