#!/bin/bash
# GPU call 20 (2 GPUs): final state -- multi-GPU parity tests and the 2-GPU bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_tests_multi_2gpu.log 2>&1
echo "multi tests rc=$?"; tail -3 gpurun_out/r2_tests_multi_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/call20_bench_m1_n2.json 2> gpurun_out/call20_bench_m1_n2.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/call20_bench_m1_n2.json') if l.startswith('{')][-1])
print('m1 n2', d['ms_per_step'], d['stage_ms'], d['e2e']['ms'], d['e2e']['stage_ms'].get('lists_equal_device_path'), d['config']['scan_symmetric'])
PY
