#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_r02w.log 2>&1; echo "pytest exit $?"; grep -a "penumbra samples\|passed\|failed\|Error\|long horizon" gpurun_out/pytest_r02w.log | tail -8
timeout 600 python bench.py > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err; echo "bench exit $?"; cat gpurun_out/bench_r02w.json | head -c 3000
