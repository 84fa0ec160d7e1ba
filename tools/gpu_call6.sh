#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd6.log 2>&1
timeout 600 python tools/ref_compare.py c1 --match --ref-runs 1 --out c1m6 > gpurun_out/cmp_c1m6.log 2>&1
echo done
