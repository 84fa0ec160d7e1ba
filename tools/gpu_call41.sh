#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff41.log 2>&1
timeout 600 python tools/time_modes.py 2 1 > gpurun_out/time_modes41.log 2>&1
timeout 900 python -m pytest tests/test_gpu_stage_golden.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu41.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu41.log
echo done
