#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_round2.py -m gpu -x -q -s 2>&1 | grep -a "penumbra samples\|passed\|failed\|Error" | tail -6
timeout 300 python scripts/split_check.py --envs 4096,8192,16384 2>&1 | tail -3 | cut -c1-60,300-420
timeout 300 python scripts/small_probe.py --envs 4096,131072 2>&1 | tail -1
