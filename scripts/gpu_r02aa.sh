#!/bin/bash
# evidence pass for the action-bucketed throughput organisation: ncu launch list of the bench command, full capture of the step
# kernel at 131072 envs, opNav single-action launch times (which bucket is the slow one)
TAG=${1:-leo_r02f}   # (captures of this pass are named *_r02f)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1; echo "launch list exit $?"
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu leo exit $?"
timeout 300 python scripts/opnav_bucket_probe.py 2>&1 | tee gpurun_out/opnav_bucket_$TAG.txt
