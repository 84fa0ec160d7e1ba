#!/bin/bash
# round-2 GPU check S (1 GPU): per-stage, per-step comparison with the reference kernels at 640x480 (C1)
mkdir -p gpurun_out
timeout 900 python tools/stage_diff_scene.py c1 1.0 5 2 2 > gpurun_out/r2s_stage_diff_c1_race2.log 2>&1
timeout 900 python tools/stage_diff_scene.py c1 1.0 5 2 1 > gpurun_out/r2s_stage_diff_c1_race1.log 2>&1
python tools/s2_seeds.py 2 > gpurun_out/r2s_s2_seeds_mode2.txt 2>&1
grep -E "^stage|Error|error" gpurun_out/r2s_stage_diff_c1_race2.log | cut -c1-900
grep -E "^stage|Error|error" gpurun_out/r2s_stage_diff_c1_race1.log | cut -c1-900
tail -1 gpurun_out/r2s_s2_seeds_mode2.txt
