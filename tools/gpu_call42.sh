#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu42.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu42.log
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff42.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref42.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref42.log
timeout 1500 python bench.py > gpurun_out/bench42.log 2>&1; echo "rc=$?" >> gpurun_out/bench42.log
echo done
