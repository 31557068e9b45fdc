#!/bin/bash
# GPU call 15 (8 GPUs): config 3 with the memory-sized log pool.
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --workload c3 --steps 4 --warmup 2 \
    > gpurun_out/call15_bench_c3_n8.json 2> gpurun_out/call15_bench_c3_n8.err
echo "bench c3 n8 rc=$?"; tail -c 300 gpurun_out/call15_bench_c3_n8.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/call15_bench_c3_n8.json') if l.startswith('{')][-1])
    print('c3 n8', d['ms_per_step'], d['stage_ms'], d['e2e'] and d['e2e']['ms'], d['config']['scan_symmetric'], d['e2e'] and d['e2e']['stage_ms'])
except Exception as e: print("no bench json", e)
PY
