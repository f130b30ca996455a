#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_split.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_split_r02t.log
timeout 300 python scripts/split_check.py --envs 4096,8192,16384 2>&1 | tail -3 | tee gpurun_out/split_r02t.txt
