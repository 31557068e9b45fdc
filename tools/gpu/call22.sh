#!/bin/bash
# GPU call 22: CTA-pair form of the filter GEMM -- targeted parity first, then A/B timing at 1M cells, then the whole GPU suite.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter" > gpurun_out/call22_filter_tests.log 2>&1
rc=$?
tail -3 gpurun_out/call22_filter_tests.log
if [ $rc -ne 0 ]; then echo "FILTER PAIR TESTS FAILED rc=$rc"; tail -40 gpurun_out/call22_filter_tests.log; exit 1; fi
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-api-e2e --no-e2e > gpurun_out/call22_bench_pair.json 2> gpurun_out/call22_bench_pair.err
tail -c 1500 gpurun_out/call22_bench_pair.json
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-api-e2e --no-e2e --option filter_cta_pair=1 > gpurun_out/call22_bench_single.json 2> gpurun_out/call22_bench_single.err
tail -c 1500 gpurun_out/call22_bench_single.json
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/call22_tests_gpu.log 2>&1
tail -3 gpurun_out/call22_tests_gpu.log
