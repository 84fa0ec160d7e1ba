#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu11.log
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd11.log 2>&1
echo done
