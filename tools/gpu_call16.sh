#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu16.log
timeout 900 python tools/fusion_time.py > gpurun_out/fusion_time.log 2>&1
echo done
