#!/bin/bash
# round 2, GPU call 1: design inputs (cluster sizes, mma.sync rate, bulk ingress, DSMEM), tcgen05 cost table,
# today's headline on this box (default and the variant library written at the end of round 1)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt
timeout 300 scripts/ubench/ubench all > gpurun_out/r2c1_ubench.jsonl 2>&1; echo "ubench rc=$?"
timeout 120 python scripts/mma_cost.py > gpurun_out/r2c1_mma_cost.jsonl 2>&1; echo "mma_cost rc=$?"
timeout 120 python scripts/mma_cost.py --ctas 148 > gpurun_out/r2c1_mma_cost_148.jsonl 2>&1; echo "mma_cost148 rc=$?"
timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2c1_bench_default.log 2>&1; echo "bench default rc=$?"
L3_LIB_VARIANT=next timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2c1_bench_next.log 2>&1; echo "bench next rc=$?"
tail -c 600 gpurun_out/r2c1_bench_default.log; echo; tail -c 600 gpurun_out/r2c1_bench_next.log; echo
cat gpurun_out/r2c1_ubench.jsonl | head -100
