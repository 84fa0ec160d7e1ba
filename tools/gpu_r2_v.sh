#!/bin/bash
# round-2 GPU check V (1 GPU): the whole GPU suite incl. the slow gate-2 scene, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
( time DPE_SLOW_TESTS=1 timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r2v_smoke.log 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/r2v_bench.log 2> gpurun_out/r2v_bench.err
echo "bench rc=$?" >> gpurun_out/r2v_bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2v_bench_ref.log 2> gpurun_out/r2v_bench_ref.err
grep -E "passed|failed|FAILED|Error|bit-identical|direction-4" gpurun_out/r2v_pytest.log; tail -2 gpurun_out/r2v_smoke.log; tail -3 gpurun_out/r2v_bench.err; tail -c 700 gpurun_out/r2v_bench.log; echo; tail -c 600 gpurun_out/r2v_bench_ref.log
