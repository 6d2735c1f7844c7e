#!/bin/bash
# usage: cfgprof.sh <cfg> [variant]: line-level ncu capture of the full shade kernel on a BASELINE config (source-page CSV per launch)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"shade_kernel" -s 4 -c 6 -o /tmp/cfgprof -f python tools/run_cfg.py $1 $2 > /dev/null 2>&1
ncu -i /tmp/cfgprof.ncu-rep --page raw --csv > gpurun_out/cfgprof_$1_raw.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/cfgprof_$1_raw.csv")))
h = rows[0]
i_name, i_t, i_id = h.index("Kernel Name"), h.index("gpu__time_duration.sum"), h.index("ID")
best = max(rows[2:], key=lambda r: float(r[i_t].replace(",", "")))
print("heaviest launch:", best[i_id], best[i_name][:40], best[i_t])
open("/tmp/cfgprof_id", "w").write(best[i_id])
PY
ID=$(cat /tmp/cfgprof_id)
# launch-skip/count select results inside the report by index
ncu -i /tmp/cfgprof.ncu-rep --page source --csv --launch-skip $ID --launch-count 1 > gpurun_out/cfgprof_$1_source.csv 2>/dev/null
ls -la gpurun_out/cfgprof_$1_*
