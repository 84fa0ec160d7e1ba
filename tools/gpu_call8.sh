#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref8.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench8.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench8.log
timeout 1200 python tools/ref_compare.py c2 --views 12 --ref-runs 1 --match --out c2v12m8 > gpurun_out/cmp_c2v12m8.log 2>&1
echo done
