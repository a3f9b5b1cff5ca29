#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r15_dbg.log
: > $L
for op in gemm_small gemm_wo gemm_qkv gemm_w13 attn_mid attn_1b attn_8b; do
  timeout 120 python scripts/dbg_big_ops.py $op >> $L 2>&1 || echo "$op FAILED rc=$?" >> $L
done
timeout 200 python scripts/dbg_big_ops.py model 1 256 >> $L 2>&1 || echo "model 1 256 FAILED" >> $L
timeout 200 python scripts/dbg_big_ops.py model 1 2048 >> $L 2>&1 || echo "model 1 2048 FAILED" >> $L
timeout 200 python scripts/dbg_big_ops.py model 16 2048 >> $L 2>&1 || echo "model 16 2048 FAILED" >> $L
grep -v "^Traceback\|^  File\|^    " $L | cut -c1-300 | tail -30
