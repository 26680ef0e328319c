#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "bilaplacian or elasticity_values or tiled_gather_limits" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -12 gpurun_out/q_pytest.log
timeout 600 python scratch/bench_configs.py c5 2>&1 | grep -E "bilaplacian|COO" | cut -c1-900
