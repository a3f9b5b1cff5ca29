#!/bin/bash
# ncu --set full of one gate|up swapped-role GEMM (EPI 2) and one Wdown (EPI 1) launch inside an 8B batch-32 decode step
mkdir -p gpurun_out
CMD="python scripts/bench_shapes.py 8b-b32"
ncu --set full --clock-control none --import-source on -k regex:gemm_swap_kernel -s 5000 -c 4 -f -o gpurun_out/r28_swap $CMD > gpurun_out/r28_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r28_ncu.log
ncu -i gpurun_out/r28_swap.ncu-rep --page raw --csv > gpurun_out/r28_swap_raw.csv; wc -c gpurun_out/r28_swap_raw.csv
