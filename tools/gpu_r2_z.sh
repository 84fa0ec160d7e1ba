#!/bin/bash
# round-2 GPU check Z (4 GPUs): the bench line at N=4 as the driver launches it
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 3 ) > gpurun_out/r2z_bench4.log 2> gpurun_out/r2z_bench4.err
echo "rc=$?" >> gpurun_out/r2z_bench4.err
tail -3 gpurun_out/r2z_bench4.err; head -c 700 gpurun_out/r2z_bench4.log
