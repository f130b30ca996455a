#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/duo_check.py --envs 4096,8192,16384 2>&1 | tail -5 | tee gpurun_out/duo_r02q.txt
timeout 300 python scripts/duo_check.py --envs 4096 --stress 2>&1 | tail -2 | tee -a gpurun_out/duo_r02q.txt
