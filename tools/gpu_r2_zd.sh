#!/bin/bash
# round-2 GPU check ZD (1 GPU): fusion with block masks (device vs reference cloud, pipeline) + the rest of the parity file
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/r2zd_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2zd_pytest.log
grep -E "passed|failed|FAILED|Error|fusion" gpurun_out/r2zd_pytest.log | head -20
