#!/bin/bash
# round-2 GPU check Y (8 GPUs): config 5 as written (300 views 1920x1080, fusion=True) with the final kernels and the device-side fusion append
mkdir -p gpurun_out
( time timeout 800 python tools/run_config.py c5 --fusion --gpus 8 --repeat 2 --no-sidecar --no-normal ) > gpurun_out/r2y_c5_g8.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_c5_g8.log
ls -la /tmp/cfg_c5_vNone_s1.0/DPE/DPE.ply >> gpurun_out/r2y_c5_g8.log 2>&1
tail -c 2500 gpurun_out/r2y_c5_g8.log
