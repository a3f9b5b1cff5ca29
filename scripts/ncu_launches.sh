#!/bin/bash
# per-launch device times of a few decode steps at the headline shape (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --steps 1 --warmup 3 --total-len 140"
$CMD > gpurun_out/ncu_launches_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 12 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"; python3 - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches.csv')) if len(r)>10]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); 
for r in rows[1:]:
    print(r[ik][:90], r[iv])
PY
