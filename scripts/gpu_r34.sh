#!/bin/bash
mkdir -p gpurun_out
T=r34
timeout 300 python -m pytest tests/test_ragged_gpu.py -m gpu -q --timeout 120 > gpurun_out/${T}_ragged.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ragged.log
tail -25 gpurun_out/${T}_ragged.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q --timeout 120 --deselect tests/test_ragged_gpu.py > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
