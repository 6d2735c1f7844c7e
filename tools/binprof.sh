#!/bin/bash
# Line-level ncu captures of the setup / binning kernels (cfg5 bench), exported as source-page CSVs (read with
# tools/sass_profile.py against a cubin of the same build), then A/B builds given as arguments.
mkdir -p gpurun_out
for k in prim_setup fine_bin coarse_pairs; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 4 -c 1 -o /tmp/binprof_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  ncu -i /tmp/binprof_$k.ncu-rep --page source --csv > gpurun_out/binprof_${k}_source.csv 2>/dev/null
  ncu -i /tmp/binprof_$k.ncu-rep --page raw --csv > gpurun_out/binprof_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out/binprof_*
bash tools/variants_cfg.sh "$@" 2>&1 | tee gpurun_out/binprof_variants.txt
