#!/bin/bash
mkdir -p gpurun_out
for v in ${VARIANTS:-d0 dnoenv dnofsw dboth}; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs ${ENVS:-4096} 2>&1 | tail -1
done | tee gpurun_out/small_variants_r02p.txt
