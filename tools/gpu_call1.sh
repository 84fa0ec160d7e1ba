#!/bin/bash
# GPU call: golden vectors, gpu tests, bench (both arms), ncu launch list + full capture, gate-2 comparison
set -x
mkdir -p gpurun_out tests/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python oracle/make_golden.py > gpurun_out/golden.log 2>&1 && cp tests/golden/ref_probe_c1.npz gpurun_out/
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python tools/prof_cmd.py 1 > gpurun_out/prof_cmd.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python tools/prof_cmd.py 1 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_half -s 24 -c 1 -f -o gpurun_out/r01_strong_edge python tools/prof_cmd.py 1 > gpurun_out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_half -s 3 -c 1 -f -o gpurun_out/r01_strong_s0 python tools/prof_cmd.py 1 > gpurun_out/ncu_full0.log 2>&1
timeout 600 python tools/ref_compare.py c1 --seed2 --out c1s > gpurun_out/cmp_c1s.log 2>&1
timeout 1200 python tools/ref_compare.py c2 --views 12 --ref-runs 1 --seed2 --out c2v12s > gpurun_out/cmp_c2v12s.log 2>&1
echo done
