#!/bin/bash
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scratch/exch_time.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -8
