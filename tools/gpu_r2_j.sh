#!/bin/bash
# round-2 GPU check J (1 GPU): all GPU tests, smoke, bench c4 (weak-texture config)
mkdir -p gpurun_out
( time DPE_SLOW_TESTS=1 timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2j_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2j_smoke.log
( time timeout 2400 python bench.py --config c4 --steps 1 --warmup 1 ) > gpurun_out/r2j_bench_c4.log 2> gpurun_out/r2j_bench_c4.err
echo "bench rc=$?" >> gpurun_out/r2j_bench_c4.err
grep -E "passed|failed|FAILED|bit-identical|fusion vs" gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_smoke.log; tail -5 gpurun_out/r2j_bench_c4.err; tail -c 1500 gpurun_out/r2j_bench_c4.log
