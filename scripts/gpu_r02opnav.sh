#!/bin/bash
# opNav evidence pass: bench line of the opNav workload, ncu launch list, full captures of both pass kernels
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python bench.py --workload opnav --steps 6 --warmup 3 > gpurun_out/bench_opnav_$TAG.json 2> gpurun_out/bench_opnav_$TAG.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_opnav_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_opnav_$TAG.csv python bench.py --workload opnav --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_opnav_$TAG.log 2>&1; echo "launch list exit $?"
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:opnav_pass1 -s 3 -c 1 -f -o gpurun_out/prof_opnav_p1_$TAG python bench.py --workload opnav --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_opnav_p1_$TAG.log 2>&1; echo "ncu p1 exit $?"
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:opnav_pass2 -s 3 -c 1 -f -o gpurun_out/prof_opnav_p2_$TAG python bench.py --workload opnav --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_opnav_p2_$TAG.log 2>&1; echo "ncu p2 exit $?"
