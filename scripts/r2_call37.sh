#!/bin/bash
# several TMA-issuing lanes in the two tcgen05 GEMMs: parity subset, then 1 / default (swap 4, classic 2) / 4 lanes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity2_gpu.py -m gpu -x -q --timeout 600 -k "linear or forward or batched or true_width or generate" 2>&1 | tail -4 | tee gpurun_out/r37_pytest.log
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-prefill 1b 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', d['config'], 'prefill ms %.2f' % d['prefill_ms'], 'tensor frac %.3f' % d['prefill_tensor_frac'], ('decode ms %.3f hbm frac %.3f' % (d['decode_ms_per_step'], d['decode_hbm_frac'])) if 'decode_ms_per_step' in d else '')" | tee -a gpurun_out/r37_ab.log; }
run L3_LIB_VARIANT=l1
run L3_X=default
run L3_LIB_VARIANT=l4
