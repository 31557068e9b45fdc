#!/bin/bash
# GPU call 18 (2 GPUs): where a rank's scan time goes (phase clock, debug_flags bit 5).
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --option debug_flags=32 \
    > gpurun_out/call18_bench_m1_n2.json 2> gpurun_out/call18_phases_n2.err
grep "em2 sym rank" gpurun_out/call18_phases_n2.err | tail -24
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --option debug_flags=32 > gpurun_out/call18_bench_m1_n1.json 2> gpurun_out/call18_phases_n1.err
grep "em2 sym rank" gpurun_out/call18_phases_n1.err | tail -8
