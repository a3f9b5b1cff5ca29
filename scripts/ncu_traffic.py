#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` capture of decode_stack_kernel: DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), duration and the pipe / issue figures quoted in DESIGN.md 4.0.

  ncu -i gpurun_out/r02_stack.ncu-rep --page raw --csv > profiles/r02_stack_kernel_ncu_full_raw.csv
  python scripts/ncu_traffic.py profiles/r02_stack_kernel_ncu_full_raw.csv "<config text>" """
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, config = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
data = [r for r in rows[rows.index(hdr) + 2:] if len(r) == len(hdr) and "decode_stack_kernel" in r[hdr.index("Kernel Name")]]
assert data, "no decode_stack_kernel launch in the capture"
r = data[-1]
def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return v * scale
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"decode_stack_kernel": {
    "config": config, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "traffic_bytes_per_launch": int(rd + wr),
    "duration_us_under_ncu": val("gpu__time_duration.sum"),
    "source": f"{os.path.relpath(path, ROOT)} (ncu --set full --clock-control none, one launch)"}}
for k in ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"):
    if k in hdr:
        try:
            out["decode_stack_kernel"][k] = float(r[hdr.index(k)].replace(",", ""))
        except ValueError:
            pass
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=2)
print(json.dumps(out, indent=2))
