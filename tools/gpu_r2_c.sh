#!/bin/bash
# round-2 GPU check C (1 GPU): GPU tests incl. kernel-variant equality, per-class profile old vs new kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
for v in 5 0; do
  timeout 600 python tools/prof_cmd.py 3 c2 12 $v > gpurun_out/r2c_prof_c2_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_prof_c2_v$v.log
  timeout 900 python tools/prof_cmd.py 2 c4 6 $v 0.5 > gpurun_out/r2c_prof_c4_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_prof_c4_v$v.log
done
tail -c 2500 gpurun_out/r2c_pytest.log
for f in gpurun_out/r2c_prof_*.log; do echo "== $f"; tail -16 $f; done
