#!/bin/bash
mkdir -p gpurun_out
for v in r2e ta tb tc tab tabc; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 131072 --steps 20 2>&1 | tail -1
done | tee gpurun_out/big_variants_r02f.txt
