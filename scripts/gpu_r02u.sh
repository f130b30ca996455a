#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:leo_split -s 3 -c 1 -f -o gpurun_out/prof_split4096_r02c \
    python scripts/small_probe.py --envs 4096 --steps 3 --warmup 2 > gpurun_out/ncu_split4096_r02c.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_split4096_r02c.log
