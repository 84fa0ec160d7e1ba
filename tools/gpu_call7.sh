#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/prof_cmd.py 1 > gpurun_out/prof_cmd7.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_weak_list -s 8 -c 1 -f -o gpurun_out/r01d_weak python tools/prof_cmd.py 1 > gpurun_out/ncu_full7a.log 2>&1
echo done
