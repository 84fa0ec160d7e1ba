#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/ref_compare.py c4 --views 6 --scale 0.5 --ref-runs 2 --match --seed2 --out c4hv6m > gpurun_out/cmp_c4hv6m.log 2>&1
echo done
