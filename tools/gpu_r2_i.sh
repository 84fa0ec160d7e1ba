#!/bin/bash
# round-2 GPU check I (2 GPUs): multi-GPU fusion (faithful single-rank and sharded), 2-rank bench, gate-2 with the new product default
mkdir -p gpurun_out
timeout 900 python tools/run_config.py c5 --views 24 --fusion --gpus 2 --repeat 2 > gpurun_out/r2i_c5v24_g2.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_c5v24_g2.log
timeout 900 python tools/run_config.py c5 --views 24 --fusion --gpus 2 --sharded-fusion --repeat 2 > gpurun_out/r2i_c5v24_g2s.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_c5v24_g2s.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 ) > gpurun_out/r2i_bench2.log 2> gpurun_out/r2i_bench2.err
echo "bench rc=$?" >> gpurun_out/r2i_bench2.err
( time timeout 900 python -m pytest tests/test_gpu_gate2.py tests/test_gpu_parity.py -q -s ) > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log
tail -c 1200 gpurun_out/r2i_c5v24_g2.log; tail -c 1200 gpurun_out/r2i_c5v24_g2s.log; tail -c 400 gpurun_out/r2i_bench2.err; tail -c 1200 gpurun_out/r2i_bench2.log; grep -E "passed|failed|FAILED" gpurun_out/r2i_pytest.log
