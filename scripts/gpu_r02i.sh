#!/bin/bash
mkdir -p gpurun_out
for v in ong1e onz1 onz2 onz4; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python bench.py --workload opnav --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-180
done | tee gpurun_out/opnav_variants_r02i.txt
