#!/bin/bash
# ncu full capture of the step kernel for one library build.  Usage: bash scripts/gpu_profile.sh <tag> [lib.so]
TAG=$1; LIBSO=$2
mkdir -p gpurun_out
[ -n "$LIBSO" ] && export BSKENV_LIB=$PWD/$LIBSO
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json | cut -c1-400
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
