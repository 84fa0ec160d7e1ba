#!/bin/bash
# round-2 GPU check L (1 GPU): GPU tests after the fusion rewrite, C5-shape fusion timing, ncu of the anchor search
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
timeout 900 python tools/run_config.py c5 --views 24 --fusion --repeat 2 --no-normal > gpurun_out/r2l_c5v24_g1.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_c5v24_g1.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_list -s 2 -c 1 -f -o gpurun_out/r02_anchor_search python tools/prof_cmd.py 1 c4 6 0 0.5 > gpurun_out/r2l_ncu1.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2l_pytest.log; tail -c 900 gpurun_out/r2l_c5v24_g1.log; grep -E "k_list" gpurun_out/r2l_ncu1.log | head -3
