#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ref_compare.py c1 --match --ref-runs 2 --exact --out c1z > gpurun_out/cmp_c1z.log 2>&1
timeout 1200 python tools/ref_compare.py c2 --views 12 --ref-runs 2 --match --exact --out c2v12z > gpurun_out/cmp_c2v12z.log 2>&1
timeout 1200 python tools/ref_compare.py c4 --views 6 --scale 0.5 --ref-runs 2 --match --exact --out c4hv6z > gpurun_out/cmp_c4hv6z.log 2>&1
rm -f gpurun_out/cmp_*_view0.npz
echo done
