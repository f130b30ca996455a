#!/bin/bash
# One gpurun call for iteration: GPU parity tests of the default build, then time every variants/libbskenv_*.so,
# then (optional) an ncu full capture of one variant.  Usage: bash scripts/gpu_quick.sh <tag> [variant-to-profile]
TAG=${1:-q}; PROF=$2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
bash scripts/bench_variants.sh
if [ -n "$PROF" ]; then
  export BSKENV_LIB=$PWD/variants/libbskenv_${PROF}.so
  timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
      python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
fi
