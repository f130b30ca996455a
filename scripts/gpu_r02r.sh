#!/bin/bash
BSKENV_LIB=$PWD/variants/libbskenv_dprof.so timeout 300 python scripts/small_probe.py --envs 4096 --steps 1 --warmup 2 2>&1 > gpurun_out/dprof.txt
tail -135 gpurun_out/dprof.txt | grep BLK | sort -k4 -n | awk 'NR<=3 || NR>=120'
