#!/bin/bash
# usage: gpu_multi.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/multi_n$N.json 2> gpurun_out/multi_n$N.err
echo "rc=$?"; cut -c1-1800 gpurun_out/multi_n$N.json; tail -5 gpurun_out/multi_n$N.err
