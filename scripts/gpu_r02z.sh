#!/bin/bash
# round 2, second evidence pass (after the split organisation and the libm-free penumbra): GPU suite, full bench line, ncu launch
# list of the bench command, full captures of the throughput kernel (131072 envs) and of the split kernel (4096 envs)
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1; echo "launch list exit $?"
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o gpurun_out/prof_leo_$TAG python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_leo_$TAG.log 2>&1; echo "ncu leo exit $?"
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:leo_split -s 3 -c 1 -f -o gpurun_out/prof_split4096_$TAG python scripts/small_probe.py --envs 4096 --steps 2 > gpurun_out/ncu_split_$TAG.log 2>&1; echo "ncu split exit $?"
