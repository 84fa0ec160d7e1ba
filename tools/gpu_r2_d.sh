#!/bin/bash
# round-2 GPU check D (1 GPU): constant-memory stage block; occupancy variants; isolated NCC loop
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 600 python tools/prof_cmd.py 3 c2 12 0 > gpurun_out/r2d_prof_c2.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_prof_c2.log
DPE_LIB=$PWD/dpe-mvs_b200/lib/libdpe_b200_c5.so timeout 600 python tools/prof_cmd.py 3 c2 12 0 > gpurun_out/r2d_prof_c2_ctas5.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_prof_c2_ctas5.log
DPE_LIB=$PWD/dpe-mvs_b200/lib/libdpe_b200_c3.so timeout 600 python tools/prof_cmd.py 3 c2 12 0 > gpurun_out/r2d_prof_c2_ctas3.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_prof_c2_ctas3.log
timeout 600 python tools/ncc_study.py > gpurun_out/r2d_ncc_study.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_ncc_study.log
DPE_LIB=$PWD/dpe-mvs_b200/lib/libdpe_b200_c5.so timeout 600 python tools/ncc_study.py > gpurun_out/r2d_ncc_study_ctas5.log 2>&1
timeout 900 python tools/prof_cmd.py 2 c4 6 0 0.5 > gpurun_out/r2d_prof_c4.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_prof_c4.log
tail -c 1500 gpurun_out/r2d_pytest.log
for f in gpurun_out/r2d_prof_*.log gpurun_out/r2d_ncc_*.log; do echo "== $f"; tail -16 $f; done
