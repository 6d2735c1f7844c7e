#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.per_cycle_active --clock-control none -k regex:"band_cull|coarse|fine_bin|prim_setup|shade_kernel" -s 10 -c 7 --csv --log-file gpurun_out/band_l.csv python tools/band_probe.py ${1:-8} > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/band_l.csv")) if len(r) > 5]
h = rows[0]
out = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    out.setdefault((d["ID"], d["Kernel Name"].split("(")[0][-34:]), {})[d["Metric Name"].split("__")[-1][:22]] = d["Metric Value"]
for (i, k), m in out.items():
    print(i, k, m)
PY
