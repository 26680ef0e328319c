#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log; tail -3 gpurun_out/full_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; cat gpurun_out/full_bench.json | cut -c1-2500; tail -3 gpurun_out/full_bench.err
