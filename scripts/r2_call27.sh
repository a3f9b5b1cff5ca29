#!/bin/bash
# tile-contiguous weight copies for gemm_swap + hoisted RoPE table loads: full suite, A/B at 8B batch 32, per-kernel launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/r27_pytest.log
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'tok/s %.0f' % d['decode_tok_s'], 'prefill %.1f' % d['prefill_ms'])" | tee -a gpurun_out/r27_ab.log; }
run L3_X=0
run L3_SWAP_TILED=0
run L3_LIB_VARIANT=base
run L3_PDL=1
v=new
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 235 --csv --log-file gpurun_out/r27_8b_b32_launches_$v.csv python scripts/bench_shapes.py 8b-b32 > gpurun_out/r27_ncu_$v.log 2>&1
echo "== $v ncu rc=$?"
python3 - gpurun_out/r27_8b_b32_launches_$v.csv <<'PY'
import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik][:60]+' '+r[ig]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(',',''))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v[1]/1e3:9.1f} us {100*v[1]/tot:5.1f}%  x{v[0]:4d}  mean {v[1]/v[0]/1e3:7.2f} us  {k}")
print(f"total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
PY
