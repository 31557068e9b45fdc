#!/bin/bash
# GPU call 25 (2 GPUs): the multi-GPU parity tests on the final code (CTA-pair filter GEMM, host hygiene fixes).
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_tests_multi_2gpu.log 2>&1
echo "multi tests rc=$?"; tail -3 gpurun_out/r2_tests_multi_2gpu.log
