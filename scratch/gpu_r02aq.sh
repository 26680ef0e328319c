#!/bin/bash
# e2e at N GPUs with and without pinning each rank to the CPUs next to its GPU
mkdir -p gpurun_out
N=${1:-4}; T=${2:-r02aq}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
for mode in bind nobind; do
  extra=""; [ $mode = nobind ] && extra="--no-bind"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $extra > gpurun_out/${T}_bench_n${N}_$mode.json 2> gpurun_out/${T}_bench_n${N}_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_n${N}_$mode.json').read().strip().splitlines()[-1]); print('$mode', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('pcie_gbs_all_ranks'), d['e2e'].get('host_cpus_rank0'))"
done
