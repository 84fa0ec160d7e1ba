#!/bin/bash
# round-2 GPU check ZF (1 GPU): last look — GPU suite, smoke(), a short bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2zf_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2zf_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zf_smoke.log 2>&1
( time timeout 900 python bench.py --steps 1 --warmup 1 ) > gpurun_out/r2zf_bench.log 2> gpurun_out/r2zf_bench.err
echo "bench rc=$?" >> gpurun_out/r2zf_bench.err
tail -4 gpurun_out/r2zf_pytest.log; tail -1 gpurun_out/r2zf_smoke.log; tail -2 gpurun_out/r2zf_bench.err; head -c 400 gpurun_out/r2zf_bench.log
