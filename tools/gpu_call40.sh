#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu40.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu40.log
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff40.log 2>&1
timeout 600 python tools/time_modes.py 2 1 > gpurun_out/time_modes40.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden40.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden40.log
echo done
