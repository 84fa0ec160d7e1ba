#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_cur.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_cur.log
echo done
