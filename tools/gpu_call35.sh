#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/stage0_diag.py > gpurun_out/stage0_diag35.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden35.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden35.log
echo done
