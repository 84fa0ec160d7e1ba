#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu15.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke15.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke15.log
timeout 900 python oracle/make_stage_golden.py > gpurun_out/stage_golden15.log 2>&1
timeout 600 python tools/ref_compare.py c1 --match --ref-runs 2 --seed2 --out c1m15 > gpurun_out/cmp_c1m15.log 2>&1
timeout 1200 python tools/ref_compare.py c2 --views 12 --ref-runs 2 --match --out c2v12m15 > gpurun_out/cmp_c2v12m15.log 2>&1
timeout 1200 python tools/ref_compare.py c4 --views 6 --scale 0.5 --ref-runs 1 --match --out c4hv6m15 > gpurun_out/cmp_c4hv6m15.log 2>&1
echo done
