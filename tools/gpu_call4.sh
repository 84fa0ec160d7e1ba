#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
timeout 600 python tools/ncc_study.py > gpurun_out/ncc_study4.log 2>&1
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd4.log 2>&1
timeout 600 python tools/ref_compare.py c1 --match --ref-runs 1 --out c1m4 > gpurun_out/cmp_c1m4.log 2>&1
echo done
