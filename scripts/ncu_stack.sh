#!/bin/bash
# ncu --set full capture of one cluster-resident decode step (decode_stack_kernel) at the headline shape, mid context (launch 518 = step 125 of the 4th generate: position 134).
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --steps 1 --warmup 3 --total-len 140"
$CMD > gpurun_out/ncu_stack_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_stack_kernel -s 518 -c 1 -f -o gpurun_out/r02_stack $CMD > gpurun_out/ncu_stack.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_stack.log
