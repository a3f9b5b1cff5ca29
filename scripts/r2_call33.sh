#!/bin/bash
# mma.sync GQA decode attention: parity subset + A/B at 8B batch 32
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity2_gpu.py tests/test_ragged_gpu.py -m gpu -q --timeout 600 -k "attention or batched or true_width or top1 or ragged or bf16" 2>&1 | tail -12 | tee gpurun_out/r33_pytest.log
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'hbm frac %.3f' % d['decode_hbm_frac'])" | tee -a gpurun_out/r33_ab.log; }
run L3_ATTN_MMA=1
run L3_ATTN_MMA=0
run L3_ATTN_MMA=1 L3_ATTN_TARGET_CTAS=1184
