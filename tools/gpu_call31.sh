#!/bin/bash
mkdir -p gpurun_out
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden31.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden31.log
echo done
