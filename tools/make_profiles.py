"""Turns the round-end measurement batch (tools/final_measure.sh -> gpurun_out/final_*) into the committed summaries
under profiles/: launch list, ncu --set full table per kernel (cfg5) and per config (cfg2/3/4 shade + blur), the
shade kernel's executed-work summary that bench.py quotes (profiles/shade_ncu_summary.json), bench lines, per-config
table.  usage: python tools/make_profiles.py r02 [commit]"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                                                               text=True).stdout.strip()

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def raw_rows(rep):
    """Rows of an ncu raw page: from the CSV exported on the GPU box (name_raw.csv) or from the report itself."""
    as_csv = rep.replace(".ncu-rep", "_raw.csv")
    if os.path.exists(as_csv):
        raw = open(as_csv).read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return [], [], []
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        return [], [], []
    return rr[0], rr[1], rr[2:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    u = unit.lower()
    return num(v) * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1.0)


def table(fh, hh, units, r, extra_stalls=True):
    idx = {n: i for i, n in enumerate(hh)}
    want = list(WANT)
    if extra_stalls:
        want += [n for n in hh if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
    fh.write("| metric | value | unit |\n|---|---|---|\n")
    for w in want:
        if w in idx and r[idx[w]] not in ("", "n/a"):
            fv = num(r[idx[w]])
            if fv is None:
                continue
            if w.startswith("smsp__average_warps_issue_stalled_") and fv < 0.1:
                continue
            v = f"{fv:.2f}" if abs(fv) < 1000 else f"{fv:,.0f}"
            fh.write(f"| {w} | {v} | {units[idx[w]]} |\n")


# ---- launch list (cfg5, bench.py)
rows = [r for r in csv.reader(open(os.path.join(G, "final_launches.csv"))) if len(r) > 5]
h = rows[0]
per = OrderedDict()
for r in rows[1:]:
    d = dict(zip(h, r))
    per.setdefault(d["Kernel Name"].split("(")[0], []).append(float(d["Metric Value"]) / 1000.0)
ours = {k: v for k, v in per.items() if "fdc::" in k and "mip_down" not in k}
bench = json.load(open(os.path.join(G, "final_bench.json")))
frame_us = sum(sum(v) / len(v) for v in ours.values())
with open(os.path.join(P, f"{tag}_launches_cfg5_4k.md"), "w") as fh:
    fh.write(f"# ncu launch list -- cfg5 3840x2160, one B200 ({tag}, commit {commit})\n\n")
    fh.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 64 --csv python bench.py --steps 3 "
             "--warmup 3 --no-cpu-baseline`\n(the first 190 launches are atlas mip-chain uploads).  Times are cold-cache and serialised "
             "by ncu: compare SHARES with the live\nCUDA-event numbers of the same build in "
             f"`{tag}_bench_cfg5_4k.json` (shade {bench['roofline']['shade_ms']} ms, setup+binning {bench['roofline']['bin_ms']} ms "
             f"launch by launch; {bench['ms_per_step']} ms per frame as one CUDA-graph launch: shade = "
             f"{100*bench['roofline']['shade_ms']/(bench['roofline']['shade_ms']+bench['roofline']['bin_ms']):.0f} % of shade + binning).\n\n")
    fh.write("| kernel | launches captured | avg us per launch | share of one frame |\n|---|---|---|---|\n")
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
        a = sum(v) / len(v)
        fh.write(f"| `{k}` | {len(v)} | {a:.1f} | {100*a/frame_us:.1f}% |\n")
    fh.write(f"\nSum over one frame: {frame_us:.1f} us.\n")
with open(os.path.join(P, f"{tag}_launches_cfg5_4k.csv"), "w") as fh:
    fh.write(open(os.path.join(G, "final_launches.csv")).read())

# ---- full capture, cfg5
hh, units, data = raw_rows(os.path.join(G, "final_full.ncu-rep"))
idx = {n: i for i, n in enumerate(hh)}
seen = OrderedDict()
for r in data:
    seen[r[idx["Kernel Name"]].split("(")[0]] = r  # the last capture of each kernel
summary = None
with open(os.path.join(P, f"{tag}_ncu_full_kernels.md"), "w") as fh:
    fh.write(f"# ncu --set full summary ({tag}, commit {commit}), cfg5 3840x2160\n\n")
    fh.write("`ncu --set full --import-source on --clock-control none -k regex:\"shade_kernel|fine_bin|coarse_pairs|"
             "prim_setup\" -s 10 -c 5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n(read with `ncu -i ... --page raw "
             "--csv`).  ncu flushes caches before every kernel, so DRAM bytes of the binning kernels are cold-cache figures; in a frame their "
             "input was just written by the previous kernel and sits in L2.\n")
    for name, r in seen.items():
        fh.write(f"\n## {name}\n\n")
        table(fh, hh, units, r)
        if "shade_kernel" in name and num(r[idx["gpu__time_duration.sum"]]) > 50:
            rd, wr = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]), to_bytes(
                r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            summary = {"kernel": name.replace("void ", "").replace("fdc::", ""), "workload": "cfg5_4k", "commit": commit,
                       "source": f"profiles/{tag}_ncu_full_kernels.md (ncu --set full --clock-control none, one launch of bench.py --steps 2 "
                                 "--warmup 3 --no-cpu-baseline)",
                       "duration_us": num(r[idx["gpu__time_duration.sum"]]), "warp_instructions": num(r[idx["smsp__inst_executed.sum"]]),
                       "issue_active": num(r[idx["smsp__issue_active.avg.per_cycle_active"]]),
                       "pipe_fma_pct": num(r[idx["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]]),
                       "pipe_alu_pct": num(r[idx["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]]),
                       "pipe_xu_pct": num(r[idx["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]]),
                       "warps_active": num(r[idx["sm__warps_active.avg.per_cycle_active"]]),
                       "registers": num(r[idx["launch__registers_per_thread"]]),
                       "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr}
if summary:
    json.dump(summary, open(os.path.join(P, "shade_ncu_summary.json"), "w"), indent=1)

# ---- cfg2 / cfg3 / cfg4: shade and blur kernels
with open(os.path.join(P, f"{tag}_ncu_configs.md"), "w") as fh:
    fh.write(f"# ncu --set full: shade and blur kernels of cfg2, cfg3, cfg4 ({tag}, commit {commit})\n\n")
    fh.write("`ncu --set full --clock-control none -k regex:\"shade_kernel|blur_h|blur_v\" -c 24 python tools/run_cfg.py <cfg>` -- the last "
             "capture of each kernel (a frame replayed launch by launch).  Blur roofline: a pass reads and writes 4 bytes per region pixel; "
             "`hbm_frac` = (read + written algorithmic bytes) / duration / 6545 GB/s (MEASURED_PEAKS.json).\n")
    for c in (2, 3, 4):
        hh, units, data = raw_rows(os.path.join(G, f"final_cfg{c}.ncu-rep"))
        if not hh:
            continue
        idx = {n: i for i, n in enumerate(hh)}
        per_k = OrderedDict()
        for r in data:
            per_k.setdefault(r[idx["Kernel Name"]].split("(")[0], []).append(r)
        fh.write(f"\n# cfg{c}\n")
        for name, rs in per_k.items():
            # the heaviest launch of that kernel in the frame (a frame has one shade launch per segment)
            r = max(rs, key=lambda x: num(x[idx["gpu__time_duration.sum"]]) or 0.0)
            if (num(r[idx["gpu__time_duration.sum"]]) or 0) < 6.0 and "blur" not in name:
                continue  # the launch of the pair that returns at once
            fh.write(f"\n## cfg{c}: {name} ({len(rs)} launches captured, heaviest shown)\n\n")
            table(fh, hh, units, r, extra_stalls=False)
            if "blur" in name:
                grid = num(r[idx["launch__grid_size"]])
                dur = num(r[idx["gpu__time_duration.sum"]])
                rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
                wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
                fh.write(f"\nDRAM traffic {(rd + wr) / 1e6:.2f} MB in {dur:.1f} us = {(rd + wr) / (dur * 1e-6) / 1e9:.0f} GB/s "
                         f"= {(rd + wr) / (dur * 1e-6) / 1e9 / 6545.3:.3f} of the measured HBM copy bandwidth ({int(grid)} CTAs: the panel is "
                         "small, the pass is launch- and latency-bound, not bandwidth-bound).\n")

# ---- bench lines and configs
json.dump(bench, open(os.path.join(P, f"{tag}_bench_cfg5_4k.json"), "w"), indent=1)
json.dump(json.load(open(os.path.join(G, "final_bench_reference.json"))), open(os.path.join(P, f"{tag}_bench_reference_cpu.json"), "w"), indent=1)
cfg = [json.loads(l) for l in open(os.path.join(G, "final_configs.txt")) if l.startswith("{")]
json.dump(cfg, open(os.path.join(P, f"{tag}_configs_all.json"), "w"), indent=1)
ss = os.path.join(G, "final_shade_stats.json")
if os.path.exists(ss):
    lines = [l for l in open(ss) if l.startswith("{")]
    if lines:
        json.dump(json.loads(lines[-1]), open(os.path.join(P, f"{tag}_shade_visits.json"), "w"), indent=1)
print("wrote profiles for", tag, "commit", commit, "frame_us", round(frame_us, 1), "summary", summary)
