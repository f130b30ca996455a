#!/bin/bash
# One gpurun call for the opNav path: bench line, ncu launch list + one full capture of opnav_step_kernel.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_check_opnav.sh [tag]
TAG=${1:-opnav_r01}
OUT=gpurun_out
mkdir -p $OUT
echo "== bench --workload opnav" ; timeout 900 python bench.py --workload opnav --steps 5 --warmup 3 --cpu-seconds 6 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench exit $?"; cat $OUT/bench_${TAG}.json; tail -5 $OUT/bench_${TAG}.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --workload opnav --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_${TAG}.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full capture of the opNav step kernel"
timeout 1200 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:opnav_step -s 3 -c 1 -f -o $OUT/prof_${TAG} \
    python bench.py --workload opnav --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
ls -la $OUT | tail -8
