#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stack_gpu.py tests/test_packed_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest stack+packed rc=$?"; tail -15 gpurun_out/r2c6_pytest.log
timeout 600 python scripts/stack_determinism.py 256 200 6 24 2>&1 | tail -1
timeout 900 python scripts/stack_sweep.py '{}' '{"L3_STACK_PF":64,"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_PF":128,"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_PF":256,"L3_STACK_KV_EVICT_FIRST":1}' '{"L3_STACK_KV_EVICT_FIRST":1}' | tee gpurun_out/r2c6_sweep.jsonl
timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c6_timeline_pf0.txt 2>&1; cat gpurun_out/r2c6_timeline_pf0.txt
L3_STACK_PF=128 L3_STACK_KV_EVICT_FIRST=1 timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c6_timeline_pf128.txt 2>&1; head -9 gpurun_out/r2c6_timeline_pf128.txt; tail -12 gpurun_out/r2c6_timeline_pf128.txt
