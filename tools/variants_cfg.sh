#!/bin/bash
# usage: tools/variants_cfg.sh "<flags A>" "<flags B>" ... -- per variant: gpu/shade ms of every BASELINE config (same box)
for v in "$@"; do
  FDC_NVCC_EXTRA="$v" python figdraw_b200/build.py --force > /dev/null 2>&1
  echo "== variant: $v"
  FDC_SKIP_ORACLE=1 python tools/bench_configs.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  %-44s gpu %.4f graph %.4f shade %.4f bin %.4f' % (d['config'], d['gpu_ms'], d.get('graph_ms', 0), d['shade_ms'], d['bin_ms']))"
done
