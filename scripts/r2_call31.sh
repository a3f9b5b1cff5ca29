#!/bin/bash
# L2 prefetch of Wo / leading gate|up k-blocks during the batched decode attention: parity subset + sweep at 8B batch 32
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity2_gpu.py -m gpu -x -q --timeout 600 -k "batched or true_width or attention_decode" 2>&1 | tail -3
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'hbm frac %.3f' % d['decode_hbm_frac'])" | tee -a gpurun_out/r31_ab.log; }
run L3_ATTN_PF_WO=0 L3_ATTN_PF_W13_KB=0
run L3_ATTN_PF_WO=1 L3_ATTN_PF_W13_KB=0
run L3_ATTN_PF_WO=1 L3_ATTN_PF_W13_KB=8
run L3_ATTN_PF_WO=1 L3_ATTN_PF_W13_KB=16
run L3_ATTN_PF_WO=1 L3_ATTN_PF_W13_KB=24
run L3_ATTN_PF_WO=0 L3_ATTN_PF_W13_KB=16
