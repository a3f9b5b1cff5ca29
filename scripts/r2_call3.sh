#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/r2c3_pytest_stack.log 2>&1; echo "pytest stack rc=$?"; tail -15 gpurun_out/r2c3_pytest_stack.log
timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2c3_bench.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2c3_bench.log
