#!/bin/bash
# round-2 GPU check ZB (1 GPU): the load phase after making the nvJPEG handle lazy; JPEG-facing GPU tests
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py 1 3 c2 > gpurun_out/r2zb_e2e.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2zb_pytest.log 2>&1
grep -o '"seconds": [0-9.]*\|"load_s": [0-9.]*\|"ctx_s": [0-9.]*\|"upload_s": [0-9.]*\|"stages_s": [0-9.]*' gpurun_out/r2zb_e2e.log | paste - - - - -
tail -2 gpurun_out/r2zb_pytest.log
