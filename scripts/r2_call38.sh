#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 300 -k "batched or swapped" 2>&1 | tail -2
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 8b-b1 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', d['config'], 'prefill ms %.2f' % d['prefill_ms'], 'decode ms %.3f hbm frac %.3f' % (d['decode_ms_per_step'], d['decode_hbm_frac']))" | tee -a gpurun_out/r38_ab.log; }
run L3_PDL_NORM=1
run L3_PDL_NORM=0
run L3_PDL_NORM=1
