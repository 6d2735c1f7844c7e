#!/bin/bash
# usage: tools/variants.sh "<nvcc extra flags A>" "<flags B>" ...  -- builds each variant ON THE GPU BOX and runs the bench
for v in "$@"; do
  FDC_NVCC_EXTRA="$v" python figdraw_b200/build.py --force > /dev/null 2>&1
  echo "== variant: $v"
  grep -A2 "shade_kernel" figdraw_b200/csrc/build.log | grep -E "registers|spill" | head -2
  python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e_ms', d['e2e']['ms_per_step'])"
done
