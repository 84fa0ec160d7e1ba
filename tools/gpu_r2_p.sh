#!/bin/bash
# round-2 GPU check P (1 GPU): where the first fine sweep differs; fp16 top-level texture (exactness + speed)
mkdir -p gpurun_out
python tools/s2_seeds.py 2 > gpurun_out/r2p_s2_seeds_mode2.txt 2>&1
python tools/s2_seeds.py 1 > gpurun_out/r2p_s2_seeds_mode1.txt 2>&1
DPE_TEX_F16=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage_golden.py -m gpu -q -s > gpurun_out/r2p_pytest_f16.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_pytest_f16.log
timeout 600 python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2p_prof_f32.log 2>&1
DPE_TEX_F16=1 timeout 600 python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2p_prof_f16.log 2>&1
tail -5 gpurun_out/r2p_pytest_f16.log; grep -E "strong_sweep|classify|weak_sweep|wall" gpurun_out/r2p_prof_f32.log gpurun_out/r2p_prof_f16.log; tail -4 gpurun_out/r2p_s2_seeds_mode2.txt
