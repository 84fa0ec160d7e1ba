#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu38.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu38.log
timeout 300 python tools/stage0_diag.py > gpurun_out/stage0_diag38.log 2>&1
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff38.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden38.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden38.log
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd38.log 2>&1
DPE_ARITH=1 timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd38_fast.log 2>&1
echo done
