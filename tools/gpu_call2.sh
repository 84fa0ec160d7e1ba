#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/tex_study.py > gpurun_out/tex_study.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench2.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench2.log
timeout 600 python tools/prof_cmd.py 1 > gpurun_out/prof_cmd2.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches2.csv python tools/prof_cmd.py 1 > gpurun_out/ncu_launches2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_half -s 40 -c 1 -f -o gpurun_out/r01b_strong_edge python tools/prof_cmd.py 1 > gpurun_out/ncu_full2a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_full -s 11 -c 1 -f -o gpurun_out/r01b_classify python tools/prof_cmd.py 1 > gpurun_out/ncu_full2b.log 2>&1
echo done
