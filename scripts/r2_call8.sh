#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest stack rc=$?"; tail -5 gpurun_out/r2c8_pytest.log
timeout 600 python scripts/stack_determinism.py 256 150 6 24 2>&1 | tail -1
timeout 900 python scripts/stack_sweep.py '{}' '{"L3_LIB_VARIANT":"nb2"}' '{"L3_LIB_VARIANT":"nb1"}' '{"L3_STACK_PF":128}' '{"L3_STACK_PF":256}' '{"L3_STACK_PF":0}' '{"L3_LIB_VARIANT":"nb1","L3_STACK_PF":128}' | tee gpurun_out/r2c8_sweep.jsonl
timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c8_timeline.txt 2>&1; cat gpurun_out/r2c8_timeline.txt
L3_STACK_PF=128 timeout 300 python scripts/stack_timeline.py --len 134 > gpurun_out/r2c8_timeline_pf128.txt 2>&1; head -9 gpurun_out/r2c8_timeline_pf128.txt; tail -12 gpurun_out/r2c8_timeline_pf128.txt
