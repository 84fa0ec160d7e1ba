#!/bin/bash
# round-2 GPU check A (1 GPU): GPU tests, smoke, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
( time timeout 1200 python bench.py --steps 1 --warmup 1 ) > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_pytest.log; tail -c 600 gpurun_out/r2a_smoke.log; tail -c 1500 gpurun_out/r2a_bench.err; tail -c 2500 gpurun_out/r2a_bench.log
