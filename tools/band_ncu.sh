#!/bin/bash
# usage: band_ncu.sh [n_ranks]: stall / cache metrics of one band's shade kernel alone on one GPU
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.per_cycle_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,launch__grid_size,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:"shade_kernel" -s 8 -c 2 --csv --log-file gpurun_out/band_ncu.csv python tools/band_probe.py ${1:-8} > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/band_ncu.csv")) if len(r) > 5]
h = rows[0]
out = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    out.setdefault((d["ID"], d["Kernel Name"].split("(")[0][-36:]), {})[d["Metric Name"]] = d["Metric Value"]
for (i, k), m in out.items():
    print(i, k)
    for a, b in m.items():
        print("     %-90s %s" % (a, b))
PY
