#!/bin/bash
# round 2, session a: GPU tests + north-star bench (C4 default) + reference arm
mkdir -p gpurun_out
T=${1:-r02a}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -5 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; cut -c1-1500 gpurun_out/${T}_bench_n1.json; tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2>> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_reference_arm.json
nvidia-smi --query-gpu=name,memory.total --format=csv
