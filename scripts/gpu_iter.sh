#!/bin/bash
# One gpurun call per kernel iteration: GPU parity tests, a short bench, optionally an ncu full capture.
# Usage: bash scripts/gpu_iter.sh <tag> [ncu]
TAG=${1:-it}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}.json"))
    print("${TAG}", "ms/step %.2f"%d["ms_per_step"], "steps/s %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "b4096 ms %.2f"%d["batch4096"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "checksum", d.get("checksum"))
except Exception as e:
    print("bench FAILED", e); print(open("gpurun_out/bench_${TAG}.err").read()[-800:])
PY
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:leo_step -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
      python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
fi
