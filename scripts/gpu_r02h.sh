#!/bin/bash
mkdir -p gpurun_out
for v in kd0 kd0m2; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096,131072 --steps 20 2>&1 | tail -1 | cut -c1-400
done | tee gpurun_out/big_variants_r02h.txt
for v in ong1 ong2 ong3 ong4; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python bench.py --workload opnav --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-180
done | tee gpurun_out/opnav_variants_r02h.txt
