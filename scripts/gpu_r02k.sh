#!/bin/bash
mkdir -p gpurun_out
for v in rare b160r200; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096,131072 --steps 20 2>&1 | tail -1 | cut -c1-420
done | tee gpurun_out/big_variants_r02k.txt
