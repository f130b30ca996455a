#!/bin/bash
mkdir -p gpurun_out
for v in m1 m1l16 m1l8; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 2048,4096,8192 2>&1 | tail -1
done | tee gpurun_out/small_variants_r02c.txt
