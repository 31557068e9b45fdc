#!/bin/bash
# GPU call 24: sanity after the host-side hygiene fixes (bounce-buffer set-up, tensor-map entry point lookup): smoke + the staging / driver / filter tests.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/call24_smoke.log 2>&1; tail -1 gpurun_out/call24_smoke.log
timeout 200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_host_layer.py -x -q -m gpu -k "pageable or driver or filter or host or whole" > gpurun_out/call24_tests.log 2>&1; tail -2 gpurun_out/call24_tests.log
