#!/bin/bash
# N=2 sanity of the final build: the driver's own launch line
mkdir -p gpurun_out
T=${1:-r02bm}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "rc=$?"
tail -c 600 gpurun_out/${T}_bench_n2.json; tail -3 gpurun_out/${T}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_ref_n2.json 2> gpurun_out/${T}_ref_n2.err; echo "rc=$?"
tail -c 400 gpurun_out/${T}_ref_n2.json
