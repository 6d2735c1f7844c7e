#!/bin/bash
# per-kernel durations (ncu, cold caches, serialised) of one cfg5 frame + instruction counts of the binning kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.per_cycle_active --clock-control none -k regex:"prim_setup|coarse|fine_bin|shade_kernel" -s 12 -c 12 --csv --log-file gpurun_out/launch_times.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/launch_times.csv")) if len(r) > 5]
h = rows[0]
out = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    out.setdefault((d["ID"], d["Kernel Name"].split("(")[0][-40:]), {})[d["Metric Name"]] = d["Metric Value"]
for (i, k), m in out.items():
    print("%4s %-42s %8s us %12s inst  issue %5s  warps %6s" % (i, k, m.get("gpu__time_duration.sum"), m.get("smsp__inst_executed.sum"), m.get("smsp__issue_active.avg.per_cycle_active"), m.get("sm__warps_active.avg.per_cycle_active")))
PY
