#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py tests/test_stack_gpu.py -m gpu -x -q --timeout 800 2>&1 | tail -3
timeout 600 python scripts/stack_sweep.py '{}' | tee gpurun_out/r2c18_sweep.jsonl
timeout 600 python scripts/stack_determinism.py 256 100 6 24 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/bench_tp.py --batches 1,32 --decode 128 > gpurun_out/r2c18_tp2.log 2>&1; echo "bench_tp rc=$?"; grep '^{' gpurun_out/r2c18_tp2.log | tee gpurun_out/r2c18_tp2.jsonl | cut -c1-330
timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c18_timeline.txt 2>&1; head -9 gpurun_out/r2c18_timeline.txt; tail -1 gpurun_out/r2c18_timeline.txt
