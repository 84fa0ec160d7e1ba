#!/bin/bash
# round-2 GPU check X (8 GPUs): the bench line at N=8 with the final kernels
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2x_bench8.log 2> gpurun_out/r2x_bench8.err
echo "rc=$?" >> gpurun_out/r2x_bench8.err
tail -3 gpurun_out/r2x_bench8.err; tail -c 1500 gpurun_out/r2x_bench8.log
