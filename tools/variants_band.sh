#!/bin/bash
# usage: tools/variants_band.sh "<flags A>" ... -- per variant: one band of an 8-way partition alone on one GPU (4K and 8K)
for v in "$@"; do
  FDC_NVCC_EXTRA="$v" python figdraw_b200/build.py --force > /dev/null 2>&1
  echo "== variant: $v"
  python tools/band_probe.py 8 2>/dev/null | grep "rank 4"
  python tools/band_probe.py 8 7680 4320 2.0 2>/dev/null | grep "rank 4"
  python tools/band_probe.py 2 2>/dev/null | grep "rank 1"
done
