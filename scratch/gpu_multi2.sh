#!/bin/bash
# strong-scaling bench on N GPUs (C4 box cut in N z-slabs)
mkdir -p gpurun_out
N=${1:-2}; T=${2:-r02m}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
cut -c1-1800 gpurun_out/${T}_bench_n$N.json; tail -5 gpurun_out/${T}_bench_n$N.err
