#!/bin/bash
# End-of-round check on one B200: GPU suite, smoke, headline bench (both arms), shape sweep.
mkdir -p gpurun_out
T=${1:-final}
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 300 python bench.py > gpurun_out/${T}_bench.log 2>&1; tail -c 400 gpurun_out/${T}_bench.log; echo
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.log 2>&1; tail -c 300 gpurun_out/${T}_bench_ref.log; echo
timeout 500 python scripts/bench_shapes.py > gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
