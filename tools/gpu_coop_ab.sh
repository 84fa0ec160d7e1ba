#!/bin/bash
# GPU check of the warp-cooperative sweep + classifier (1 GPU, ~4 min): A/B against the per-pixel forms on a resident
# 12-view C2-shape scene (speed + bit-identity at full size), the whole-schedule identity test and the golden stage
# tests with the cooperative forms selected, then one ncu --set full capture of the cooperative classifier.
# The scene is pre-rendered into .ab_scene/ (git-ignored, travels with the snapshot); without it the tools render it.
mkdir -p gpurun_out
export DPE_BENCH_DIR=$PWD/.ab_scene
timeout 170 python tools/ab_coop.py c2 12 1.0 2 > gpurun_out/coop_ab.txt 2>&1; echo "rc=$?" >> gpurun_out/coop_ab.txt
tail -4 gpurun_out/coop_ab.txt
DPE_VARIANTS=0 timeout 200 python -m pytest "tests/test_gpu_parity.py::test_cooperative_scoring_gives_identical_maps" tests/test_gpu_stage_golden.py -x -q -m gpu -s > gpurun_out/coop_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/coop_pytest.txt
tail -12 gpurun_out/coop_pytest.txt
timeout 150 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_fullILi2 -s 5 -c 1 -f -o gpurun_out/r02_classify_coop python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/coop_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/coop_ncu.log
tail -3 gpurun_out/coop_ncu.log
echo done
