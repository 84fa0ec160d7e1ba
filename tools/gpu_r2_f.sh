#!/bin/bash
# round-2 GPU check F (2 GPUs): all GPU tests, pipeline with fusion on 1 and 2 GPUs (C5-shape, reduced), e2e timing on 2 GPUs
mkdir -p gpurun_out
( time DPE_SLOW_TESTS=1 timeout 1800 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
timeout 900 python tools/run_config.py c5 --views 24 --fusion > gpurun_out/r2f_c5v24_g1.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_c5v24_g1.log
timeout 900 python tools/run_config.py c5 --views 24 --fusion --gpus 2 > gpurun_out/r2f_c5v24_g2.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_c5v24_g2.log
DPE_TRACE=1 timeout 600 python tools/e2e_breakdown.py 2 3 > gpurun_out/r2f_e2e2.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_e2e2.log
grep -v "^$" gpurun_out/r2f_pytest.log | tail -40
tail -c 1500 gpurun_out/r2f_c5v24_g1.log; tail -c 1500 gpurun_out/r2f_c5v24_g2.log; grep seconds gpurun_out/r2f_e2e2.log
