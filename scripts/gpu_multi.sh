#!/bin/bash
# bench.py on N GPUs of one box, launched as the driver launches it.  Usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N TAG'
N=$1; TAG=${2:-r02c}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "exit $?"
tail -1 gpurun_out/bench_${N}gpu_$TAG.json | cut -c1-330
