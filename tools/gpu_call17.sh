#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/run_config.py c5 --views 40 --fusion > gpurun_out/cfg_c5.log 2>&1
timeout 600 python tools/prof_cmd.py 1 > gpurun_out/prof_cmd17.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches17.csv python tools/prof_cmd.py 1 > gpurun_out/ncu_launches17.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_full -s 11 -c 1 -f -o gpurun_out/r01e_classify python tools/prof_cmd.py 1 > gpurun_out/ncu17a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_half -s 30 -c 1 -f -o gpurun_out/r01e_strong_edge python tools/prof_cmd.py 1 > gpurun_out/ncu17b.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_weak_list -s 8 -c 1 -f -o gpurun_out/r01e_weak python tools/prof_cmd.py 1 > gpurun_out/ncu17c.log 2>&1
echo done
