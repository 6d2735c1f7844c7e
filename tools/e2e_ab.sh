#!/bin/bash
# e2e A/B on one box: ring depth and read-back mode of the pipelined loop (bench.py env knobs)
for r in ${RINGS:-2 3 4}; do
FDC_E2E_RING=$r python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ring=$r', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'latency', d['e2e']['latency_ms'], 'present', d['e2e_present']['ms_per_step'])"
done
FDC_E2E_ASYNC_READ=1 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('async read, ring=3', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
