#!/bin/bash
# round-2 GPU check T (1 GPU): per-stage, per-step comparison at 640x480 for the LAST view (whose scratch stays readable), with
# the reference run twice per stage
mkdir -p gpurun_out
DPE_STAGE_DIFF_REFREF=1 timeout 1200 python tools/stage_diff_scene.py c1 1.0 5 4 2 > gpurun_out/r2t_stage_diff_c1_race2.log 2>&1
DPE_STAGE_DIFF_REFREF=1 timeout 1200 python tools/stage_diff_scene.py c1 1.0 5 4 1 > gpurun_out/r2t_stage_diff_c1_race1.log 2>&1
grep -E "^stage [4-7]|Error|error" gpurun_out/r2t_stage_diff_c1_race2.log | cut -c1-1500
