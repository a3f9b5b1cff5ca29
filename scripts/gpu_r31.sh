#!/bin/bash
mkdir -p gpurun_out
T=r31
timeout 300 python scripts/bench_shapes.py 8b-32k 1b-prefill > gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
CMD="python scripts/bench_shapes.py 1b-prefill"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gemm_tc|attn_prefill_tc|rmsnorm|linear_rows|argmax|embed" -s 150 -c 120 --csv --log-file gpurun_out/${T}_1b_prefill.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
CMD="python scripts/bench_shapes.py 8b-prefill"
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gemm_tc|attn_prefill_tc|rmsnorm|linear_rows|argmax|embed" -s 300 -c 240 --csv --log-file gpurun_out/${T}_8b_prefill.csv $CMD > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/${T}_*
