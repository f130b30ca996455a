#!/bin/bash
# compute-sanitizer passes over both step kernels at small sizes (memcheck, racecheck, initcheck); run on the GPU box:
#   gpurun -- 'bash scripts/sanitize.sh'        -> gpurun_out/sanitizer_*.log
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/san_workload.py > $OUT/sanitizer_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^leo|^opnav" $OUT/sanitizer_$tool.log | cut -c1-160 | tail -4
done
echo "== compute-sanitizer --tool synccheck (split organisation: named barriers)"
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python scripts/san_split.py > $OUT/sanitizer_synccheck.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|^leo" $OUT/sanitizer_synccheck.log | cut -c1-160 | tail -4
