#!/bin/bash
# round-2 GPU check N (1 GPU): GPU tests + gate 2, default bench line, ncu launch list of the bench command, ncu --set full of the classifier
mkdir -p gpurun_out
( time DPE_SLOW_TESTS=1 timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/r2n_bench.log 2> gpurun_out/r2n_bench.err
echo "bench rc=$?" >> gpurun_out/r2n_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2n_ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_full -s 9 -c 1 -f -o gpurun_out/r02_classify python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2n_ncu_classify.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_half -s 40 -c 1 -f -o gpurun_out/r02_strong_fine python tools/prof_cmd.py 1 c2 12 0 > gpurun_out/r2n_ncu_strong.log 2>&1
grep -E "passed|failed|FAILED|Error|bit-identical" gpurun_out/r2n_pytest.log; tail -3 gpurun_out/r2n_bench.err; tail -c 600 gpurun_out/r2n_bench.log; wc -l gpurun_out/r02_launches.csv; ls -la gpurun_out/r02_classify.ncu-rep gpurun_out/r02_strong_fine.ncu-rep
