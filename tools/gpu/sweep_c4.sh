#!/bin/bash
# BASELINE.json configs[3]: LSH bit-width sweep 256/1024/4096 at 1M cells, XOR/POPC vs +-1 int8 tcgen05 Hamming variants.
# One bench line per (bits, variant) into gpurun_out/r2_c4_sweep.jsonl.
mkdir -p gpurun_out
: > gpurun_out/r2_c4_sweep.jsonl
for L in 256 1024 4096; do
  for V in mma popc; do
    steps=3; [ "$V" = popc ] && steps=1
    timeout 900 python bench.py --workload c4 --lsh $L --variant $V --steps $steps --warmup 1 --no-cpu-baseline --no-e2e \
        >> gpurun_out/r2_c4_sweep.jsonl 2>> gpurun_out/r2_c4_sweep.err
    echo "c4 L=$L $V rc=$?"
  done
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_c4_sweep.jsonl'):
    d = json.loads(ln); r = d['roofline']
    print(d['config']['lsh'], d['config']['variant'], 'sym', d['config']['scan_symmetric'], round(d['ms_per_step'], 1), 'ms', f"{d['value']:.3e}", 'pairs/s',
          r['bound'], round(r['frac'], 3), round(r['executed_frac'], 3), d['clocks'].get('sm_mhz'), d['clocks'].get('samples'), d['clocks'].get('reasons'))
PY
