#!/bin/bash
# round-2 GPU check ZC (1 GPU): ncu launch lists of the final kernels — the bench command (first 600 launches) and one whole
# view (all 8 stages) of the profile scene
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/r2zc_bench_plain.log 2>&1
echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2zc_ncu_bench.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_one_view.csv python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2zc_ncu_view.log 2>&1
wc -l gpurun_out/r02_launches.csv gpurun_out/r02_launches_one_view.csv
