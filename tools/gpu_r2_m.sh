#!/bin/bash
# round-2 GPU check M (1 GPU): GPU tests (viz, smaller anchor-search stack), anchor-search residency variants on the C4-shape scene
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
timeout 900 python tools/prof_cmd.py 2 c4 6 0 0.5 > gpurun_out/r2m_prof_c4_nb16.log 2>&1
for t in nb8 nb4 nb3; do
  DPE_LIB=$PWD/dpe-mvs_b200/lib/libdpe_b200_$t.so timeout 900 python tools/prof_cmd.py 2 c4 6 0 0.5 > gpurun_out/r2m_prof_c4_$t.log 2>&1
done
grep -E "passed|failed|FAILED|Error" gpurun_out/r2m_pytest.log
for f in gpurun_out/r2m_prof_c4_*.log; do echo "== $f"; grep -E "gen_neigh|fit_plane|wall" $f; done
