#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r18_dbg.log
: > $L
echo "== bench_shapes 1b debug sync" >> $L
L3_DEBUG_SYNC=1 timeout 300 python scripts/bench_shapes.py 1b >> $L 2>&1
echo "== bench_shapes 1b debug sync, no attn tc" >> $L
L3_ATTN_TC=0 L3_DEBUG_SYNC=1 timeout 300 python scripts/bench_shapes.py 1b >> $L 2>&1
echo "== bench_shapes 1b plain x2" >> $L
timeout 300 python scripts/bench_shapes.py 1b >> $L 2>&1
timeout 300 python scripts/bench_shapes.py 1b >> $L 2>&1
grep -v "^Traceback\|^  File\|^    " $L | cut -c1-500 | tail -40
