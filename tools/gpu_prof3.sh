#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_cmd.py 3 > gpurun_out/prof_cmd6.log 2>&1
echo done
