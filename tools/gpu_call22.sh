#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/prof_cmd.py 1 c4 6 > gpurun_out/prof_c4b.log 2>&1 && \
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_light -s 66 -c 1 -f -o gpurun_out/r01f_genneigh python tools/prof_cmd.py 1 c4 6 > gpurun_out/ncu22.log 2>&1
echo done
