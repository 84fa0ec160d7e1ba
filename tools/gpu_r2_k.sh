#!/bin/bash
# round-2 GPU check K (8 GPUs): the 8-rank bench (value + in-process e2e through dpe_mvs), and BASELINE.json configs[4] as written:
# 300 views 1920x1080 on 8 B200 with fusion=True -> DPE.ply
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2k_smi.txt 2>&1
df -h /tmp > gpurun_out/r2k_df.txt 2>&1; free -g >> gpurun_out/r2k_df.txt 2>&1; nproc >> gpurun_out/r2k_df.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 1 ) > gpurun_out/r2k_bench8.log 2> gpurun_out/r2k_bench8.err
echo "bench rc=$?" >> gpurun_out/r2k_bench8.err
( time timeout 1500 python tools/run_config.py c5 --fusion --gpus 8 --repeat 2 --no-sidecar --no-normal ) > gpurun_out/r2k_c5_g8.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_c5_g8.log
ls -la /tmp/cfg_c5_vNone_s1.0/DPE/DPE.ply >> gpurun_out/r2k_c5_g8.log 2>&1
( time timeout 900 python tools/run_config.py c5 --fusion --gpus 8 --sharded-fusion --keep-scene --no-sidecar --no-normal ) > gpurun_out/r2k_c5_g8_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_c5_g8_sharded.log
tail -c 600 gpurun_out/r2k_bench8.err; tail -c 1500 gpurun_out/r2k_bench8.log; tail -c 2500 gpurun_out/r2k_c5_g8.log; tail -c 1500 gpurun_out/r2k_c5_g8_sharded.log
