#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10.log
timeout 900 python tools/e2e_time.py > gpurun_out/e2e_time.log 2>&1
echo done
