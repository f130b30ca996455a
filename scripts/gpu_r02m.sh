#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_r02m.log 2>&1; echo "pytest exit $?"; grep -a "penumbra samples\|passed\|failed\|Error\|long horizon" gpurun_out/pytest_r02m.log | tail -8
