#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -s > gpurun_out/pytest_r02b.log 2>&1; echo "pytest exit $?"; grep -a "long horizon\|passed\|failed\|Error" gpurun_out/pytest_r02b.log | tail -8
for v in base m1 m1u m1b32 m1ub32; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096,16384 2>&1 | tail -1
done | tee gpurun_out/small_variants_r02b.txt
