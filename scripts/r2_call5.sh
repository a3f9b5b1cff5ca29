#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r2c5_pytest_stack.log 2>&1; echo "pytest stack (defaults) rc=$?"; tail -3 gpurun_out/r2c5_pytest_stack.log
L3_STACK_PF=128 L3_STACK_KV_EVICT_FIRST=1 timeout 600 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r2c5_pytest_stack_pf.log 2>&1; echo "pytest stack (pf) rc=$?"; tail -3 gpurun_out/r2c5_pytest_stack_pf.log
timeout 900 python scripts/stack_sweep.py '{}' '{"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_PF":64}' '{"L3_STACK_PF":128}' '{"L3_STACK_PF":256}' \
  '{"L3_STACK_PF":64,"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_PF":128,"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_PF":192,"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_PF":256,"L3_STACK_KV_EVICT_FIRST":1}' \
  | tee gpurun_out/r2c5_sweep.jsonl
timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c5_timeline_pf0.txt 2>&1; tail -4 gpurun_out/r2c5_timeline_pf0.txt
L3_STACK_PF=128 L3_STACK_KV_EVICT_FIRST=1 timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c5_timeline_pf128.txt 2>&1; head -9 gpurun_out/r2c5_timeline_pf128.txt; tail -1 gpurun_out/r2c5_timeline_pf128.txt
