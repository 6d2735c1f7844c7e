#!/bin/bash
# Round-end measurement batch on one B200: tests, bench lines (both arms), ncu launch list, ncu --set full of every
# kernel (cfg5) and of the shade / blur kernels of cfg2, cfg3, cfg4, all configs.  Output under gpurun_out/final_*;
# tools/make_profiles.py <tag> turns it into the committed summaries under profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.txt
python bench.py 2> gpurun_out/final_bench.err > gpurun_out/final_bench.json
python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/final_bench.err > gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 64 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"shade_kernel|fine_bin|coarse_pairs|prim_setup" -s 10 -c 5 -o gpurun_out/final_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# (reports are turned into raw-page CSVs here and deleted: gpurun only brings back 64 MiB)
ncu -i gpurun_out/final_full.ncu-rep --page raw --csv > gpurun_out/final_full_raw.csv 2>/dev/null
for c in 2 3 4; do
  ncu --set full --clock-control none -k regex:"shade_kernel|blur_h|blur_v" -s 2 -c 12 -o /tmp/final_cfg$c -f python tools/run_cfg.py $c > /dev/null 2>&1
  ncu -i /tmp/final_cfg$c.ncu-rep --page raw --csv > gpurun_out/final_cfg${c}_raw.csv 2>/dev/null
done
FDC_REPLAYS=20 python tools/bench_configs.py > gpurun_out/final_configs.txt 2>&1
python tools/shade_stats.py > gpurun_out/final_shade_stats.json 2>&1
cat gpurun_out/final_pytest.txt; head -c 600 gpurun_out/final_bench.json; echo; ls -la gpurun_out/ | head -40
