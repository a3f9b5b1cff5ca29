#!/bin/bash
mkdir -p gpurun_out
timeout 300 scripts/ubench/ubench ingress3 | tee gpurun_out/r2c9_ingress_lanes.jsonl
timeout 600 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -2
timeout 900 python scripts/stack_sweep.py '{}' '{"L3_LIB_VARIANT":"st16"}' '{"L3_STACK_PF":0,"L3_STACK_KV_EVICT_FIRST":0}' | tee gpurun_out/r2c9_sweep.jsonl
