#!/bin/bash
# round-2 GPU check W (1 GPU): GPU suite after the lazy fill of the candidate rows + per-class timing of the profile scene
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
timeout 600 python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2w_prof_c2.log 2>&1
grep -E "passed|failed|FAILED|Error|bit-identical|direction-4" gpurun_out/r2w_pytest.log
grep -E "strong_sweep|classify|weak_sweep|wall" gpurun_out/r2w_prof_c2.log
