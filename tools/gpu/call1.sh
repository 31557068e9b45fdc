#!/bin/bash
# GPU call 1 of round 2: the metric's 1M-cell workload through bench.py, launch list, ncu --set full of the paced
# symmetric sweep, measured int8 peaks.
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi --query-gpu=name,memory.total,power.limit --format=csv; } > gpurun_out/box.txt 2>&1
expressionmatrix2_b200/build/mma_peak > gpurun_out/r2_mma_peak.json 2> gpurun_out/mma_peak.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_m1_n1.json 2> gpurun_out/bench_m1_n1.err
tail -c 600 gpurun_out/bench_m1_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_m1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_m1_under_ncu.json 2> gpurun_out/ncu_launch.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scanMmaSymKernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2_scan_sym_m1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -20
