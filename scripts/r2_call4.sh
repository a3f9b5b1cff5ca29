#!/bin/bash
mkdir -p gpurun_out
python scripts/stack_determinism.py 256 300 6 24 2>&1 | tail -1
timeout 900 python -m pytest tests/test_stack_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r2c4_pytest_stack.log 2>&1; echo "pytest stack rc=$?"; tail -5 gpurun_out/r2c4_pytest_stack.log
timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2c4_bench.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2c4_bench.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2c4_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/r2c4_pytest_all.log
