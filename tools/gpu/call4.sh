#!/bin/bash
# GPU call 4: where the time goes after the near window + leader grouping (launch lists, ncu full of the far sweep), config tests.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_bench_m1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call4_m1_under_ncu.json 2> gpurun_out/call4_ncu_launch.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_bench_c2.csv \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call4_c2_under_ncu.json 2> gpurun_out/call4_ncu_launch_c2.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2b_scan_sym_far_m1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/call4_ncu_full.log 2>&1
tail -3 gpurun_out/call4_ncu_full.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 0 --launch-count 1 \
    -o gpurun_out/r2b_scan_sym_near_m1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/call4_ncu_full_near.log 2>&1
tail -3 gpurun_out/call4_ncu_full_near.log
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/call4_tests_configs.log 2>&1
echo "config tests rc=$?"; tail -5 gpurun_out/call4_tests_configs.log
