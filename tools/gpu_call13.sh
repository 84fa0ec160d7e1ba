#!/bin/bash
mkdir -p gpurun_out tests/golden
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden.log 2>&1; echo "rc=$?" >> gpurun_out/stage_golden.log
cp tests/golden/ref_stage_first.npz tests/golden/ref_stage_weak.npz gpurun_out/ 2>/dev/null
ls -la tests/golden >> gpurun_out/stage_golden.log
echo done
