#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'hbm frac %.3f' % d['decode_hbm_frac'])" | tee -a gpurun_out/r36_ab.log; }
run L3_SWAP_KSPLIT=8
run L3_SWAP_KSPLIT=2
run L3_SWAP_KSPLIT=1
run L3_SWAP_RESID_ATOMIC=0
