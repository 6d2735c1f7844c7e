#!/bin/bash
# usage: tools/ncu_cfg.sh <cfg number> -- ncu launch list (per-kernel times) for one replay of a BASELINE config
cat > /tmp/run_cfg.py <<PY
import sys; sys.path.insert(0, ".")
from figdraw_b200 import scenes_synth as ss
from figdraw_b200.cuda_context import CudaContext, render_trace
tr = ss.config_trace($1)
ctx = CudaContext(atlasSize=tr.atlas_size)
render_trace(tr, ctx)
for _ in range(3):
    ctx.replayFrame(); ctx.sync()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_cfg$1.csv python /tmp/run_cfg.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_cfg$1.csv")) if len(r)>5]
h=rows[0]
last=[dict(zip(h,r)) for r in rows[1:]]
# keep the last replay: find last occurrence of prim_setup sequence start
names=[x["Kernel Name"].split("(")[0] for x in last]
n_per=None
idx=[i for i,n in enumerate(names) if "prim_setup" in n]
segs=len(idx)//4 if len(idx)>=4 else 1
start=idx[-segs] if idx else 0
for x in last[start:]:
    print("%-60s %8.1f us" % (x["Kernel Name"].split("(")[0][-60:], float(x["Metric Value"])/1000))
PY
