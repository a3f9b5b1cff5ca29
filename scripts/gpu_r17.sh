#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r17_dbg.log
: > $L
for v in "1 256" "1 2048" "2 2048 2306" "2 2048 2114 argmax" "2 2048 2306 argmax" "16 2048 2306 argmax"; do
  echo "== model $v" >> $L
  timeout 200 python scripts/dbg_big_ops.py model $v >> $L 2>&1 || echo "FAILED" >> $L
done
timeout 300 python -m pytest tests/test_mega_gpu.py -m gpu -q --timeout 90 >> $L 2>&1
grep -v "^Traceback\|^  File\|^    " $L | cut -c1-300 | tail -40
