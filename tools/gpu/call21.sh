#!/bin/bash
# GPU call 21: ncu --set full of the second and third kernels of the 1M-cell step (filter GEMM, merge).
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sigFilterKernel --launch-skip 3 --launch-count 1 \
    -o gpurun_out/r2_sigfilter_m1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/call21_ncu1.log 2>&1
tail -1 gpurun_out/call21_ncu1.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:finalizeSymKernel --launch-skip 0 --launch-count 1 \
    -o gpurun_out/r2_finalize_sym_m1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/call21_ncu2.log 2>&1
tail -1 gpurun_out/call21_ncu2.log
