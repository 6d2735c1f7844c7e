#!/bin/bash
# usage: tools/multi_run.sh N  -- multi-GPU test + bench lines (balanced and equal bands) on N GPUs of one box
N=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/multi_test_n$N.txt
cat gpurun_out/multi_test_n$N.txt
for b in balanced equal; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --bands $b > gpurun_out/bench_n${N}_$b.json 2> gpurun_out/bench_n${N}_$b.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n${N}_$b.json") if l.startswith("{")][-1])
    x = d.get("cfg5_8k", {})
    print("N=$N bands=$b 4K ms", d["ms_per_step"], "shade", d["roofline"]["shade_ms"], "bin", d["roofline"]["bin_ms"], "e2e", d["e2e"]["ms_per_step"], "| 8K ms", x.get("ms_per_step"), "eff", x.get("efficiency"), "shade", x.get("shade_ms"), "bin", x.get("bin_ms"), "e2e", x.get("e2e_ms_per_step"), "ok", d.get("gathered_frame_equals_single_gpu"), x.get("gathered_frame_equals_single_gpu"))
    print("   bounds 4K", d.get("band_tile_rows"), "8K", x.get("band_tile_rows"))
except Exception as e:
    print("N=$N bands=$b failed:", e); print(open("gpurun_out/bench_n${N}_$b.err").read()[-1500:])
PY
done
