#!/bin/bash
# usage: multi_ring.sh N "rings..." -- e2e of the 4K workload at N GPUs for several ring depths (same box)
N=$1
for r in $2; do
  FDC_E2E_RING=$r timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-8k 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N ring=$r', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'latency', d['e2e'].get('latency_ms'))"
done
