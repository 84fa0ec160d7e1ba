#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff30.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden30.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden30.log
echo done
