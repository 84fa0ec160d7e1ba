#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -k "weak_path or stage_matches or sharding or device_fusion" -x -q > gpurun_out/memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck.log
echo done
