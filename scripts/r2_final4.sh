#!/bin/bash
# Round-2 closing evidence: full GPU suite, smoke, the driver's bench command (both arms), launch list of an 8B batch-32 decode step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --durations=5 > gpurun_out/r02_pytest_gpu.log 2>&1 ) 2>&1 | grep real; tail -9 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r02_bench_full.log 2> gpurun_out/r02_bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -1 gpurun_out/r02_bench_full.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','launch_ms','decode_step_ms_measured','whole_step_hbm_frac','per_decode_step_ms')})
print('cpu', d['cpu_baseline'])
for r in d.get('extra', []):
    print(r.get('name'), {k: r.get(k) for k in ('value','ms_per_step','error')}, 'frac', (r.get('roofline') or {}).get('frac'), 'prefill', (r.get('prefill') or {}).get('ms'), ((r.get('prefill') or {}).get('roofline') or {}).get('frac'), 'e2e', (r.get('e2e') or {}).get('value'))
"
