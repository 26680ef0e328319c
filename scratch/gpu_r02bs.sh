#!/bin/bash
# N=8 (or N given) strong scaling on the final build: the driver's own launch line
mkdir -p gpurun_out
T=${1:-r02bs}; N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err; echo "rc=$?"
tail -c 300 gpurun_out/${T}_bench_n${N}.json; tail -3 gpurun_out/${T}_bench_n${N}.err
