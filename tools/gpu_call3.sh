#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
timeout 600 python tools/ncc_study.py > gpurun_out/ncc_study.log 2>&1
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd3.log 2>&1
timeout 600 python tools/ref_compare.py c1 --match --out c1m > gpurun_out/cmp_c1m.log 2>&1
timeout 1200 python tools/ref_compare.py c2 --views 12 --ref-runs 2 --match --out c2v12m > gpurun_out/cmp_c2v12m.log 2>&1
timeout 600 python tools/ref_compare.py c1 --ref-runs 1 --out c1j > gpurun_out/cmp_c1j.log 2>&1
echo done
