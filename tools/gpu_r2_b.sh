#!/bin/bash
# round-2 GPU check B (2 GPUs): GPU tests, e2e breakdown on 1 and 2 GPUs, a short 2-rank bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
DPE_TRACE=1 timeout 600 python tools/e2e_breakdown.py 1 2 > gpurun_out/r2b_e2e1.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_e2e1.log
DPE_TRACE=1 NCCL_DEBUG=WARN timeout 600 python tools/e2e_breakdown.py 2 3 > gpurun_out/r2b_e2e2.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_e2e2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 ) > gpurun_out/r2b_bench2.log 2> gpurun_out/r2b_bench2.err
echo "bench rc=$?" >> gpurun_out/r2b_bench2.err
tail -c 1200 gpurun_out/r2b_pytest.log; tail -c 2500 gpurun_out/r2b_e2e1.log; tail -c 3500 gpurun_out/r2b_e2e2.log; tail -c 1500 gpurun_out/r2b_bench2.err; tail -c 1500 gpurun_out/r2b_bench2.log
