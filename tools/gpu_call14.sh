#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cost_diff.py > gpurun_out/cost_diff.log 2>&1
timeout 1500 python tools/ref_compare.py c4 --views 6 --scale 0.5 --ref-runs 0 --match --out c4hv6r > gpurun_out/cmp_c4hv6r.log 2>&1
echo done
