#!/bin/bash
# round 2, first GPU pass: full GPU suite, default bench line, ncu capture of the step kernel at 4096 envs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_r02a.log
timeout 900 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_r02a.json; tail -3 gpurun_out/bench_r02a.err
timeout 300 python scripts/small_probe.py --envs 1024,4096,8192,16384,32768 > gpurun_out/small_r02a.json 2>&1; cat gpurun_out/small_r02a.json
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o gpurun_out/prof_small4096_r02a \
    python scripts/small_probe.py --envs 4096 --steps 2 > gpurun_out/ncu_small_r02a.log 2>&1; echo "ncu exit $?"
