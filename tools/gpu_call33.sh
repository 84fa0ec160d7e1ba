#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff33.log 2>&1
timeout 900 python -m pytest tests/test_gpu_stage_golden.py -m gpu -q > gpurun_out/pytest_gpu33.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu33.log
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden33.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden33.log
echo done
