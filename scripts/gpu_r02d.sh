#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02d.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_r02d.log
for v in nbase nm1; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096,16384,131072 2>&1 | tail -1
done | tee gpurun_out/small_variants_r02d.txt
