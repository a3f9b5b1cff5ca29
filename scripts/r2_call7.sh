#!/bin/bash
# 2 GPUs: tensor-parallel parity + decode latency with the flag-in-data exchange; 1-GPU stack sweep rides along on GPU 0
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_tp_gpu.py tests/test_mega_gpu.py -m gpu -x -q --timeout 800 > gpurun_out/r2c7_pytest_tp.log 2>&1; echo "pytest tp+mega rc=$?"; tail -15 gpurun_out/r2c7_pytest_tp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/bench_tp.py --batches 1 --decode 128 > gpurun_out/r2c7_tp2.log 2>&1; echo "bench_tp rc=$?"; grep '^{' gpurun_out/r2c7_tp2.log | tee gpurun_out/r2c7_tp2.jsonl | cut -c1-700; tail -3 gpurun_out/r2c7_tp2.log | cut -c1-300
timeout 600 python scripts/stack_sweep.py '{}' '{"L3_STACK_PF":0,"L3_STACK_KV_EVICT_FIRST":0}' '{"L3_STACK_PF":96}' | tee gpurun_out/r2c7_sweep.jsonl
timeout 300 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -2
