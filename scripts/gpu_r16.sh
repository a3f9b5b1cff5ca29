#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r16_dbg.log
: > $L
echo "== mega off, model 1 256" >> $L
L3_MEGA=0 timeout 200 python scripts/dbg_big_ops.py model 1 256 >> $L 2>&1 || echo "FAILED" >> $L
echo "== sanitizer, model 1 256" >> $L
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dbg_big_ops.py model 1 256 > gpurun_out/r16_sanitizer.log 2>&1
grep -E "=========|model" gpurun_out/r16_sanitizer.log | head -60 >> $L
echo "== bench_shapes 1b alone" >> $L
timeout 300 python scripts/bench_shapes.py 1b >> $L 2>&1
echo "== bench_shapes s15m then 1b" >> $L
timeout 300 python scripts/bench_shapes.py s15m-b1-f32 1b >> $L 2>&1
grep -v "^Traceback\|^  File\|^    " $L | cut -c1-400 | tail -80
