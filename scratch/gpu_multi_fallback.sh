#!/bin/bash
mkdir -p gpurun_out
AFB_P2P_DISABLE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/multi_fb.json 2> gpurun_out/multi_fb.err
echo "rc=$?"; cut -c1-900 gpurun_out/multi_fb.json; grep -v "OMP_NUM\|\*\*\*" gpurun_out/multi_fb.err | tail -5
