#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref58.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref58.log
timeout 1500 python bench.py > gpurun_out/bench58.log 2>&1; echo "rc=$?" >> gpurun_out/bench58.log
timeout 600 python tools/prof_cmd.py 1 > gpurun_out/prof_cmd58.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches58.csv python tools/prof_cmd.py 1 > gpurun_out/ncu_launches58.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_full -s 11 -c 1 -f -o gpurun_out/r01h_classify python tools/prof_cmd.py 1 > gpurun_out/ncu58a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_half -s 30 -c 1 -f -o gpurun_out/r01h_strong_edge python tools/prof_cmd.py 1 > gpurun_out/ncu58b.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_weak_list -s 8 -c 1 -f -o gpurun_out/r01h_weak python tools/prof_cmd.py 1 > gpurun_out/ncu58c.log 2>&1
echo done
