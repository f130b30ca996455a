#!/bin/bash
# Times every variants/libbskenv_*.so with a short bench run (on the GPU box).
mkdir -p gpurun_out
for f in variants/libbskenv_*.so; do
  name=$(basename $f .so)
  BSKENV_LIB=$PWD/$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$name.json"))
    print("$name", "ms/step %.2f"%d["ms_per_step"], "steps/s %.3e"%d["value"], "b4096 ms %.2f"%d["batch4096"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$name", "FAILED", e); print(open("gpurun_out/var_$name.err").read()[-500:])
PY
done
