#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_cmd.py 1 > gpurun_out/prof_cmd43.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches43.csv python tools/prof_cmd.py 1 > gpurun_out/ncu_launches43.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_full -s 11 -c 1 -f -o gpurun_out/r01g_classify python tools/prof_cmd.py 1 > gpurun_out/ncu43a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_half -s 30 -c 1 -f -o gpurun_out/r01g_strong_edge python tools/prof_cmd.py 1 > gpurun_out/ncu43b.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_weak_list -s 8 -c 1 -f -o gpurun_out/r01g_weak python tools/prof_cmd.py 1 > gpurun_out/ncu43c.log 2>&1
echo done
