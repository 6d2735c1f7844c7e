"""Join an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the same cubin and
print executed warp-instructions and stall samples per source line.
usage: sass_profile.py shade_src.csv shade.sass fdc_shade.cu [top]"""
import csv
import re
import sys
from collections import defaultdict

csv_path, sass_path, src_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]

# nvdisasm: per function, instruction offsets with the current "line N" marker (outermost inlined-at chain kept too)
funcs = []
cur = None
line = None
for l in open(sass_path):
    m = re.match(r"\.text\.(\S+):", l.strip())
    if m:
        cur = {"name": m.group(1), "ins": []}
        funcs.append(cur)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        line = int(m.group(2))
        inl = re.findall(r"line (\d+)", m.group(3))
        chain = [line] + [int(x) for x in inl]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur is not None:
        cur["ins"].append((int(m.group(1), 16), chain if line else [0], m.group(2).strip()))

# CSV rows are per function in address order; split on address discontinuities
groups = []
prev = None
for r in data:
    a = int(r[0], 16)
    if prev is None or a != prev + 16:
        groups.append([])
    groups[-1].append(r)
    prev = a
src = open(src_path).read().split("\n")
total = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"total warp instructions {total/1e6:.1f} M in {len(groups)} functions")
by_line = defaultdict(lambda: [0, 0, 0])   # inst, samples, outer-line
by_outer = defaultdict(lambda: [0, 0])
for g in groups:
    f = next((f for f in funcs if len(f["ins"]) == len(g)), None)
    if f is None:
        print("no nvdisasm function with", len(g), "instructions; have", [(f["name"][:30], len(f["ins"])) for f in funcs])
        continue
    n = sum(int(r[col["Instructions Executed"]]) for r in g)
    print(f"function {f['name'][:60]}: {len(g)} SASS, {n/1e6:.1f} M executed")
    for r, (off, chain, text) in zip(g, f["ins"]):
        ie = int(r[col["Instructions Executed"]])
        sm = int(r[col["# Samples"]])
        by_line[chain[0]][0] += ie
        by_line[chain[0]][1] += sm
        by_outer[chain[-1]][0] += ie
        by_outer[chain[-1]][1] += sm
tot_s = sum(v[1] for v in by_line.values()) or 1
print("\n-- by innermost source line")
for ln, (ie, sm, _) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {ie/1e6:8.2f} M {100*ie/total:5.1f}%  samples {100*sm/tot_s:5.1f}%  | {src[ln-1].strip()[:110] if 0 < ln <= len(src) else ''}")
print("\n-- by outermost (call-site) line")
for ln, (ie, sm) in sorted(by_outer.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{ln:5d} {ie/1e6:8.2f} M {100*ie/total:5.1f}%  samples {100*sm/tot_s:5.1f}%  | {src[ln-1].strip()[:110] if 0 < ln <= len(src) else ''}")
