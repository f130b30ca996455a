#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -a "penumbra samples\|passed\|failed\|Error\|long horizon" | tail -6
timeout 300 python scripts/split_check.py --envs 4096,16384 2>&1 | tail -2 | cut -c1-60,300-420
timeout 300 python scripts/small_probe.py --envs 4096,65536,131072 2>&1 | tail -1
