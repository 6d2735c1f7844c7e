"""Turns the round-end measurement batch (tools/final_measure.sh -> gpurun_out/) into the committed summaries under
profiles/: launch list, ncu --set full table per kernel, shade-kernel DRAM traffic, bench lines, per-config table.
usage: python tools/make_profiles.py r01"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, "final_launches.csv"))) if len(r) > 5]
h = rows[0]
per = OrderedDict()
for r in rows[1:]:
    d = dict(zip(h, r))
    per.setdefault(d["Kernel Name"].split("(")[0], []).append(float(d["Metric Value"]) / 1000.0)
ours = {k: v for k, v in per.items() if "fdc::" in k}
bench = json.load(open(os.path.join(G, "final_bench.json")))
frame_us = sum(sum(v) / len(v) for v in ours.values() if "mip_down" not in "".join(ours.keys()) or True)
frame_us = sum(sum(v) / len(v) for k, v in ours.items() if "mip_down" not in k)
with open(os.path.join(P, f"{tag}_launches_cfg5_4k.md"), "w") as fh:
    fh.write(f"# ncu launch list -- cfg5 3840x2160, one B200 ({tag}, final kernels)\n\n")
    fh.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 48 --csv python bench.py --steps 3 "
             "--warmup 3 --no-cpu-baseline`\n(the first 190 launches are atlas mip-chain uploads). Times are cold-cache and serialised "
             "by ncu: compare SHARES with the live\nCUDA-event numbers of the same build in "
             f"`{tag}_bench_cfg5_4k.json` (shade {bench['roofline']['shade_ms']} ms, setup+binning {bench['roofline']['bin_ms']} ms "
             f"of a {bench['ms_per_step']} ms frame = {100*bench['roofline']['shade_ms']/bench['ms_per_step']:.0f} % / "
             f"{100*bench['roofline']['bin_ms']/bench['ms_per_step']:.0f} %).\n\n")
    fh.write("| kernel | launches captured | avg us per launch | share of one frame |\n|---|---|---|---|\n")
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
        if "mip_down" in k:
            continue
        a = sum(v) / len(v)
        fh.write(f"| `{k}` | {len(v)} | {a:.1f} | {100*a/frame_us:.1f}% |\n")
    fh.write(f"\nSum over one frame: {frame_us:.1f} us.\n")

# ---- full capture
rep = os.path.join(G, "final_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh = rr[0]
units = rr[1]
idx = {n: i for i, n in enumerate(hh)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]
want += [n for n in hh if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
seen = {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    seen[name] = r  # keep the last capture of each kernel
traffic = None
with open(os.path.join(P, f"{tag}_ncu_full_kernels.md"), "w") as fh:
    fh.write(f"# ncu --set full summary ({tag}, final kernels), cfg5 3840x2160\n\n")
    fh.write("`ncu --set full --import-source on --clock-control none -k regex:\"shade_kernel|fine_bin|coarse_bin|prim_setup|coarse_scan\" "
             "-s 12 -c 6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n(read with `ncu -i ... --page raw --csv`). "
             "ncu flushes caches before every kernel, so DRAM bytes of the binning kernels are cold-cache figures; in a frame their "
             "input was just written by the previous kernel and sits in L2.\n")
    for name, r in seen.items():
        fh.write(f"\n## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            if w in idx and r[idx[w]] not in ("", "n/a"):
                v = r[idx[w]]
                try:
                    fv = float(v.replace(",", ""))
                    if w.startswith("smsp__average_warps_issue_stalled_") and fv < 0.1:
                        continue
                    v = f"{fv:.2f}" if abs(fv) < 1000 else f"{fv:,.0f}"
                except ValueError:
                    pass
                fh.write(f"| {w} | {v} | {units[idx[w]]} |\n")
        if "shade_kernel" in name:
            def mb(x):
                v = float(r[idx[x]].replace(",", ""))
                u = units[idx[x]].lower()
                return v * (1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1e9 if u.startswith("gbyte") else 1.0)
            traffic = {"kernel": "shade_kernel", "dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_write": mb("dram__bytes_write.sum")}
            traffic["traffic_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
            traffic["source"] = f"profiles/{tag}_ncu_full_kernels.md (ncu --set full, one launch, cfg5 3840x2160)"
if traffic:
    json.dump(traffic, open(os.path.join(P, f"{tag}_shade_traffic.json"), "w"), indent=1)

# ---- bench lines and configs
json.dump(bench, open(os.path.join(P, f"{tag}_bench_cfg5_4k.json"), "w"), indent=1)
json.dump(json.load(open(os.path.join(G, "final_bench_reference.json"))), open(os.path.join(P, f"{tag}_bench_reference_cpu.json"), "w"), indent=1)
cfg = [json.loads(l) for l in open(os.path.join(G, "final_configs.txt")) if l.startswith("{")]
json.dump(cfg, open(os.path.join(P, f"{tag}_configs_all.json"), "w"), indent=1)
print("wrote profiles for", tag, "frame_us", round(frame_us, 1), "traffic", traffic)
