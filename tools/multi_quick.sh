#!/bin/bash
# usage: multi_quick.sh N [extra bench args] -- GPU test suite (incl. the multi-GPU test) + one bench line on N GPUs
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n${N}_final.json") if l.startswith("{")][-1])
x = d.get("cfg5_8k", {})
print("N=$N 4K ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "latency", d["e2e"].get("latency_ms"), "| 8K ms", x.get("ms_per_step"), "eff", x.get("efficiency"), "e2e", x.get("e2e_ms_per_step"), "latency", x.get("e2e_latency_ms"), "ok", d.get("gathered_frame_equals_single_gpu"), x.get("gathered_frame_equals_single_gpu"))
PY
