#!/bin/bash
mkdir -p gpurun_out
T=r29
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.log 2>&1; tail -c 600 gpurun_out/${T}_ref.log
