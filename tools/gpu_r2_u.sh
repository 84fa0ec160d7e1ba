#!/bin/bash
# round-2 GPU check U (1 GPU): seed pixels of the first sweep of stage 5 at 640x480
mkdir -p gpurun_out
timeout 900 python tools/sweep_seeds_scene.py c1 1.0 5 5 2 > gpurun_out/r2u_seeds_c1_stage5.txt 2>&1
timeout 900 python tools/sweep_seeds_scene.py c1 1.0 5 6 2 > gpurun_out/r2u_seeds_c1_stage6.txt 2>&1
head -50 gpurun_out/r2u_seeds_c1_stage5.txt | cut -c1-400
