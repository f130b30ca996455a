#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02g.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_r02g.log
timeout 600 python bench.py --workload opnav --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_opnav_r02g.json 2> gpurun_out/bench_opnav_r02g.err; cut -c1-300 gpurun_out/bench_opnav_r02g.json
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:opnav_step -s 3 -c 1 -f -o gpurun_out/prof_opnav_r02g \
    python bench.py --workload opnav --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_opnav_r02g.log 2>&1; echo "ncu exit $?"
