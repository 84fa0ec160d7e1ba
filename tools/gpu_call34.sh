#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/stage0_diag.py > gpurun_out/stage0_diag.log 2>&1
echo done
