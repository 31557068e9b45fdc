#!/bin/bash
# GPU call 8: config-4 sweep at 1M cells; symmetric vs one-directional at c2; API-level e2e at 1M cells.
bash tools/gpu/sweep_c4.sh
for flag in --symmetric --one-directional; do
timeout 600 python bench.py --workload c2 $flag --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/call8_bench_c2$flag.json 2> gpurun_out/call8_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/call8_bench_c2$flag.json')); print('c2 $flag', d['ms_per_step'], d['stage_ms'], d['config']['scan_symmetric'])"
done
timeout 900 python tools/e2e_host.py --workload m1 --repeat 2 > gpurun_out/r2_e2e_host_m1.json 2> gpurun_out/call8_e2e_host.err
echo "e2e_host m1 rc=$?"; cut -c1-1200 gpurun_out/r2_e2e_host_m1.json; tail -c 300 gpurun_out/call8_e2e_host.err
