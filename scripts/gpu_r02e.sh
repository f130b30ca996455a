#!/bin/bash
mkdir -p gpurun_out
for v in r2e r2eu; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096,16384,131072 2>&1 | tail -1
done | tee gpurun_out/small_variants_r02e.txt
