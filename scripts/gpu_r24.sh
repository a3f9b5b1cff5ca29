#!/bin/bash
# ncu evidence (1 GPU): launch list of the headline bench, full capture of its dominant kernel,
# DRAM traffic of the persistent decode kernel and of the prefill kernels at the 8B shape.
mkdir -p gpurun_out
T=r24
CMD="python bench.py --steps 1 --warmup 1 --total-len 140 --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 1500 -c 3 -o gpurun_out/${T}_attn_decode $CMD > gpurun_out/${T}_ncu2.log 2>&1
CMD2="python scripts/bench_shapes.py 8b-b1"
$CMD2 > gpurun_out/${T}_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"decode_mega|gemm_tc|attn_prefill_tc" -c 200 --csv --log-file gpurun_out/${T}_8b_metrics.csv $CMD2 > gpurun_out/${T}_ncu3.log 2>&1
ls -la gpurun_out/${T}_*; tail -3 gpurun_out/${T}_ncu1.log gpurun_out/${T}_ncu2.log gpurun_out/${T}_ncu3.log
