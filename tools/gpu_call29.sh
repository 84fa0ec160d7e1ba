#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff29.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden29.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden29.log
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd29.log 2>&1
echo done
