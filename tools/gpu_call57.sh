#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu57.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu57.log
timeout 600 python tools/ref_compare.py c1 --match --ref-runs 2 --exact --out c1w > gpurun_out/cmp_c1w.log 2>&1
timeout 1200 python tools/ref_compare.py c2 --views 12 --ref-runs 1 --match --exact --out c2v12w > gpurun_out/cmp_c2v12w.log 2>&1
timeout 1200 python tools/ref_compare.py c4 --views 6 --scale 0.5 --ref-runs 1 --match --exact --out c4hv6w > gpurun_out/cmp_c4hv6w.log 2>&1
rm -f gpurun_out/cmp_*_view0.npz
echo done
