#!/bin/bash
# opNav three-kernel interval: variants of the noise / dynamics kernels (variants/libbskenv_*.so), probe + per-kernel times
mkdir -p gpurun_out
for v in nb5 nb6 nb6d2 nb6u4; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/opnav_split_probe.py --envs ${OPNAV_N:-113664,75776} 2>&1 | tr '\n' ' '; echo
done | tee gpurun_out/opnav_split_variants.txt
