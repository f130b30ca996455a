#!/bin/bash
# One gpurun call: GPU test-suite, smoke, bench, ncu launch list + one full capture of the step kernel.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv > $OUT/gpu_${TAG}.txt 2>&1
nproc >> $OUT/gpu_${TAG}.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_${TAG}.log; tail -15 $OUT/pytest_gpu_${TAG}.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_${TAG}.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke_${TAG}.log
echo "== bench" ; timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench exit $?"; cat $OUT/bench_${TAG}.json; tail -5 $OUT/bench_${TAG}.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err; cat $OUT/bench_ref_${TAG}.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $OUT/ncu_launch_${TAG}.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full capture of the step kernel"
timeout 1200 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o $OUT/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
echo "== opNav path"; bash scripts/gpu_check_opnav.sh opnav_${TAG}
