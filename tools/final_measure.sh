#!/bin/bash
# Round-end measurement batch on one B200: tests, bench line, ncu launch list, ncu full capture, all configs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.txt
python bench.py 2> gpurun_out/final_bench.err > gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/final_bench.err > gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 48 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"shade_kernel|fine_bin|coarse_bin|prim_setup|coarse_scan" -s 12 -c 6 -o gpurun_out/final_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
FDC_REPLAYS=40 python tools/bench_configs.py > gpurun_out/final_configs.txt 2>&1
cat gpurun_out/final_pytest.txt; head -c 600 gpurun_out/final_bench.json; echo; ls -la gpurun_out/
