#!/bin/bash
# round-2 GPU check O (1 GPU): direction-4 snapshot mode — stage replay + gate 2 (C1 and the C4-shape scene)
mkdir -p gpurun_out
( time DPE_SLOW_TESTS=1 timeout 1500 python -m pytest tests/test_gpu_stage_golden.py tests/test_gpu_gate2.py -m gpu -q -s ) > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
grep -E "passed|failed|FAILED|Error|bit-identical" gpurun_out/r2o_pytest.log
