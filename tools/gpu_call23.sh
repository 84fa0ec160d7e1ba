#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu23.log
timeout 900 python tools/prof_cmd.py 2 c4 6 > gpurun_out/prof_c4c.log 2>&1
echo done
