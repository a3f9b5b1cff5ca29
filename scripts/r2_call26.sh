#!/bin/bash
# per-kernel launch list of an 8B batch-32 decode step: balanced K-ranges (default library) against the previous schedule (base)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 600 -k "attention" 2>&1 | tail -3
for v in "" base; do
  L3_LIB_VARIANT=$v ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 235 --csv --log-file gpurun_out/r26_8b_b32_launches_${v:-new}.csv python scripts/bench_shapes.py 8b-b32 > gpurun_out/r26_ncu_${v:-new}.log 2>&1
  echo "== ${v:-new} ncu rc=$?"
  python3 - gpurun_out/r26_8b_b32_launches_${v:-new}.csv <<'PY'
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
done
