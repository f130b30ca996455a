#!/bin/bash
mkdir -p gpurun_out
BSKENV_LIB=$PWD/variants/libbskenv_sprof.so timeout 300 python scripts/small_probe.py --envs 4096 --steps 1 --warmup 2 2>&1 > gpurun_out/sprof.txt
tail -135 gpurun_out/sprof.txt | grep BLK | sort -k4 -n | awk 'NR<=3 || NR==64 || NR>=124'
tail -1 gpurun_out/sprof.txt | cut -c1-90
