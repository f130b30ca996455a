#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02n.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_r02n.log
bash scripts/sanitize.sh
