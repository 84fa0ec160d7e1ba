#!/bin/bash
# round-2 GPU check G (1 GPU): all GPU tests, K2/K3 golden, ncu of the anchor search and the weak sweep on the C4-shape scene
mkdir -p gpurun_out
timeout 600 python oracle/make_k2k3_golden.py > gpurun_out/r2g_k2k3.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_k2k3.log
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 900 python tools/prof_cmd.py 2 c4 6 0 0.5 > gpurun_out/r2g_prof_c4.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_prof_c4.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_listILi5 -s 3 -c 1 -f -o gpurun_out/r02_gen_neighbours python tools/prof_cmd.py 1 c4 6 0 0.5 > gpurun_out/r2g_ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_weak_list -s 14 -c 1 -f -o gpurun_out/r02_weak_sweep python tools/prof_cmd.py 1 c4 6 0 0.5 > gpurun_out/r2g_ncu2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_halfILi1 -s 14 -c 1 -f -o gpurun_out/r02_strong_edge python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2g_ncu3.log 2>&1
tail -3 gpurun_out/r2g_k2k3.log; grep -v "^$" gpurun_out/r2g_pytest.log | grep -E "passed|failed|FAILED|Error|fusion vs" | head; tail -14 gpurun_out/r2g_prof_c4.log; ls -la gpurun_out/*.ncu-rep
