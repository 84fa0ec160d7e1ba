#!/bin/bash
# round-2 GPU check ZA (2 GPUs): the reference arm launched under torchrun like the driver does at N > 1 (rank 0 alone works)
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > gpurun_out/r2za_ref2.log 2> gpurun_out/r2za_ref2.err
echo "rc=$?" >> gpurun_out/r2za_ref2.err
tail -3 gpurun_out/r2za_ref2.err; head -c 500 gpurun_out/r2za_ref2.log
