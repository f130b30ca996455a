#!/bin/bash
# Build the CUDA library to a cubin, print ptxas figures and the loop structure of the reference-config step kernel.
# Usage: bash scripts/sass_loops.sh [extra nvcc flags]
cd "$(dirname "$0")/../basilisk_env_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -cubin "$@" -o /tmp/bskenv.cubin csrc/bskenv.cu 2>&1 | grep -A2 "leo_step_kernelILi3ELb0ELb1" | tail -2
cuobjdump -sass /tmp/bskenv.cubin | awk '/Function : .*leo_step_kernelILi3ELb0ELb1/{f=1} f{print} /Function : /{if(f&&!/leo_step_kernelILi3ELb0ELb1/)exit}' > /tmp/k.sass
python /root/repo/scripts/loops.py /tmp/k.sass | sort -k5 -n | tail -6
