#!/bin/bash
mkdir -p gpurun_out
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden24.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden24.log
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd24.log 2>&1
echo done
