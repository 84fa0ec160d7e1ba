#!/bin/bash
# round-2 GPU check Q (1 GPU): live direction-4 reads in the reference's launch geometry — stage replay, gate 2
mkdir -p gpurun_out
python tools/s2_seeds.py 1 > gpurun_out/r2q_s2_seeds_mode1.txt 2>&1
( time DPE_SLOW_TESTS=1 timeout 1500 python -m pytest tests/test_gpu_stage_golden.py tests/test_gpu_gate2.py tests/test_gpu_parity.py -m gpu -q -s ) > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
grep -E "passed|failed|FAILED|Error|bit-identical|direction-4" gpurun_out/r2q_pytest.log; head -1 gpurun_out/r2q_s2_seeds_mode1.txt; tail -2 gpurun_out/r2q_s2_seeds_mode1.txt
