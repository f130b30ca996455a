#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_duo.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_duo_r02t.log
timeout 600 python scripts/duo_check.py --envs 4096,8192,16384 2>&1 | tail -3 | tee gpurun_out/duo_r02t.txt
