#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
timeout 600 python scratch/bench_configs.py > gpurun_out/configs.log 2>&1; cat gpurun_out/configs.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -20 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --mode replicate > gpurun_out/bench_n2_rep.json 2> gpurun_out/bench_n2_rep.err; cat gpurun_out/bench_n2_rep.json; tail -5 gpurun_out/bench_n2_rep.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
