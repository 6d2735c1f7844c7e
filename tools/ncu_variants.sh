#!/bin/bash
# usage: tools/ncu_variants.sh "<flags A>" "<flags B>" ... -- per variant: shade_kernel duration / instructions / issue utilisation under ncu
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct
K=${NCU_KERNEL:-shade_kernel}
mkdir -p gpurun_out
i=0
for v in "$@"; do
  FDC_NVCC_EXTRA="$v" python figdraw_b200/build.py --force > /dev/null 2>&1
  echo "== variant: $v"
  ncu --metrics $M --clock-control none -k regex:$K -s 4 -c 1 --csv --log-file gpurun_out/ncu_var_$i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/ncu_var_$i.csv")) if len(r) > 5]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print("  %-80s %s %s" % (d.get("Metric Name"), d.get("Metric Value"), d.get("Metric Unit")))
PY
  i=$((i+1))
done
