#!/bin/bash
# round-2 GPU check ZE (1 GPU): seed pixels of the first sweep of stage 5 at 1120x840, where the edge-adaptive pass (up to 22
# steps) and the fixed pass (11 steps) of the sampling differ
mkdir -p gpurun_out
timeout 800 python tools/sweep_seeds_scene.py c2 0.7 5 5 2 > gpurun_out/r2ze_seeds_c2_stage5.txt 2>&1
head -30 gpurun_out/r2ze_seeds_c2_stage5.txt | cut -c1-420
rm -f gpurun_out/sweep_seeds_c2_*.npz
