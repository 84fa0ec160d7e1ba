#!/bin/bash
# round-2 GPU check E (1 GPU): gate-2 test (reference build vs ours, C1 + C4-shape), short bench with the reference sample leg, reference arm
mkdir -p gpurun_out
( time DPE_SLOW_TESTS=1 timeout 1500 python -m pytest tests/test_gpu_gate2.py -x -q -s ) > gpurun_out/r2e_gate2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_gate2.log
( time timeout 1200 python bench.py --steps 1 --warmup 1 ) > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err
echo "bench rc=$?" >> gpurun_out/r2e_bench.err
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2e_bench_ref.log 2> gpurun_out/r2e_bench_ref.err
echo "ref rc=$?" >> gpurun_out/r2e_bench_ref.err
tail -c 3000 gpurun_out/r2e_gate2.log; tail -c 800 gpurun_out/r2e_bench.err; tail -c 2000 gpurun_out/r2e_bench.log; tail -c 600 gpurun_out/r2e_bench_ref.err; tail -c 2500 gpurun_out/r2e_bench_ref.log
