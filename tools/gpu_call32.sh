#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage_golden.py -m gpu -q > gpurun_out/pytest_gpu32.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu32.log
echo done
