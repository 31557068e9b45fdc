#!/bin/bash
# GPU call 7 (2 GPUs): multi-GPU parity tests (in-library driver + one process per GPU), 2-GPU bench of the metric's workload.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/call7_gpus.txt; nvidia-smi topo -m >> gpurun_out/call7_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_tests_multi_2gpu.log 2>&1
echo "multi tests rc=$?"; tail -15 gpurun_out/r2_tests_multi_2gpu.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter or golden or streamed or row_blocks" > gpurun_out/call7_tests_kernels.log 2>&1
echo "kernel tests rc=$?"; tail -3 gpurun_out/call7_tests_kernels.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 \
    > gpurun_out/call7_bench_m1_n2.json 2> gpurun_out/call7_bench_m1_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/call7_bench_m1_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call7_bench_m1_n2.json') if l.startswith('{')][-1])
    print('m1 n2', d['ms_per_step'], d['stage_ms'], d['e2e'], d['roofline']['frac'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload c2 --steps 5 --warmup 3 \
    > gpurun_out/call7_bench_c2_n2.json 2> gpurun_out/call7_bench_c2_n2.err
echo "bench c2 n2 rc=$?"; tail -c 600 gpurun_out/call7_bench_c2_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call7_bench_c2_n2.json') if l.startswith('{')][-1])
    print('c2 n2', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['config']['scan_symmetric'])
except Exception as e: print("no bench json", e)
PY
