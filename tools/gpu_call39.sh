#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_modes.py 2 1 > gpurun_out/time_modes39.log 2>&1
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd39.log 2>&1
DPE_ARITH=1 timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd39_fast.log 2>&1
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden39.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden39.log
timeout 900 python -m pytest tests/test_gpu_stage_golden.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu39.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu39.log
echo done
