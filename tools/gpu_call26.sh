#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu26.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu26.log
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff26.log 2>&1
echo done
