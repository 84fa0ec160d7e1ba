#!/bin/bash
# round-2 GPU check H (1 GPU): GPU tests (exactness prints), weak-sweep occupancy variants, ncu of anchor search + strong sweep
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
for t in w8 w5; do
  DPE_LIB=$PWD/dpe-mvs_b200/lib/libdpe_b200_$t.so timeout 900 python tools/prof_cmd.py 2 c4 6 0 0.5 > gpurun_out/r2h_prof_c4_$t.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_prof_c4_$t.log
done
timeout 900 python tools/prof_cmd.py 2 c4 6 0 0.5 > gpurun_out/r2h_prof_c4.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_prof_c4.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_list -s 6 -c 1 -f -o gpurun_out/r02_gen_neighbours python tools/prof_cmd.py 1 c4 6 0 0.5 > gpurun_out/r2h_ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_half -s 14 -c 1 -f -o gpurun_out/r02_strong_edge python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2h_ncu3.log 2>&1
grep -v "^$" gpurun_out/r2h_pytest.log | grep -E "passed|failed|FAILED|Error|fusion vs|bit-identical|assert" | head -20
for f in gpurun_out/r2h_prof_c4*.log; do echo "== $f"; grep -E "weak_sweep|gen_neigh|wall" $f; done; ls -la gpurun_out/r02_*.ncu-rep; grep -E "k_list|k_half" gpurun_out/r2h_ncu1.log | head -5
