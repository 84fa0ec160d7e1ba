#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu25.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu25.log
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff25.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden25.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden25.log
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd25.log 2>&1
DPE_ARITH=2 timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd25_exact.log 2>&1
echo done
