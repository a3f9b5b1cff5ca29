#!/bin/bash
# balanced K-ranges in gemm_swap + staged decode attention + deep pre-dependency weight prefetch: parity, then A/B at 8B batch 32
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/r25_pytest.log
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'tok/s %.0f' % d['decode_tok_s'], 'prefill %.1f' % d['prefill_ms'])" | tee -a gpurun_out/r25_ab.log; }
run L3_LIB_VARIANT=base
run L3_X=0
run L3_ATTN_STAGED=0
run L3_SWAP_BALANCED=0
run L3_PDL=1
run L3_PDL=1 L3_LIB_VARIANT=base
run L3_ATTN_TARGET_CTAS=1184
