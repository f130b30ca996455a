#!/bin/bash
mkdir -p gpurun_out
for v in u2; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096,131072 --steps 20 2>&1 | tail -1 | cut -c1-420
done | tee gpurun_out/big_variants_r02l.txt
for v in pk2 pk3g1 pk3g2 pk3g4; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python bench.py --workload opnav --opnav-envs 113664 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-180
done | tee gpurun_out/opnav_variants_r02l.txt
