#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/bench_shapes.py 8b-b32"
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 235 --csv --log-file gpurun_out/r02_8b_b32_launches.csv $CMD > gpurun_out/ncu_8bb32.log 2>&1
echo "ncu rc=$?"; python3 - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_8b_b32_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik][:70]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(',',''))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v[1]/1e3:9.1f} us {100*v[1]/tot:5.1f}%  x{v[0]:4d}  mean {v[1]/v[0]/1e3:7.2f} us  {k}")
print(f"total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
PY
