#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1 2; do
  echo "== L2PF=$pf" >> gpurun_out/r12_gemv.log
  L3_GEMV_L2PF=$pf timeout 120 python scripts/gemv_sweep.py 8b 1b >> gpurun_out/r12_gemv.log 2>&1
done
for env in "X=1" "L3_PDL=1" "L3_PDL=1 L3_GEMV_L2PF=1" "L3_GEMV_L2PF=1"; do
  echo "== $env" >> gpurun_out/r12_shapes.log
  env $env timeout 200 python scripts/bench_shapes.py 1b 8b-b1 >> gpurun_out/r12_shapes.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_1b.csv python scripts/bench_shapes.py 1b > gpurun_out/r12_ncu.log 2>&1
tail -40 gpurun_out/r12_gemv.log; python scripts/show_shapes.py gpurun_out/r12_shapes.log
