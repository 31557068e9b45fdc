#!/bin/bash
# GPU call 23: ncu --set full of the CTA-pair filter GEMM at 1M cells.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sigFilterKernel --launch-skip 3 --launch-count 1 \
    -o gpurun_out/r2g_sigfilter_pair_m1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/call23_ncu.log 2>&1
tail -2 gpurun_out/call23_ncu.log | cut -c1-300
