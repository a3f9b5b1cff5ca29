#!/bin/bash
# One GPU-box pass: tensor-core unit tests first (bounded, a hung mbarrier wait must not eat
# the lease), then the whole GPU suite, smoke and a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-run}
timeout 150 python -m pytest tests -m gpu -q -x --timeout 60 -k "tcgen05" > gpurun_out/${TAG}_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/${TAG}_tc.log
tail -15 gpurun_out/${TAG}_tc.log
if grep -q "tc rc=0" gpurun_out/${TAG}_tc.log; then
  timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -15 gpurun_out/${TAG}_pytest.log
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
  tail -3 gpurun_out/${TAG}_smoke.log
  timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
  tail -3 gpurun_out/${TAG}_bench.log
fi
