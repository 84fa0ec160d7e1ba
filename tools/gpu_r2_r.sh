#!/bin/bash
# round-2 GPU check R (1 GPU): reference geometry with direction 4 last; batched cost loads of the position search
mkdir -p gpurun_out
python tools/s2_seeds.py 1 > gpurun_out/r2r_s2_seeds_mode1.txt 2>&1
( time DPE_SLOW_TESTS=1 timeout 1500 python -m pytest tests/test_gpu_stage_golden.py tests/test_gpu_gate2.py tests/test_gpu_parity.py -m gpu -q -s ) > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 600 python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2r_prof_c2.log 2>&1
grep -E "passed|failed|FAILED|Error|bit-identical|direction-4" gpurun_out/r2r_pytest.log; head -1 gpurun_out/r2r_s2_seeds_mode1.txt; tail -2 gpurun_out/r2r_s2_seeds_mode1.txt
grep -E "strong_sweep|classify|weak_sweep|wall" gpurun_out/r2r_prof_c2.log
