#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cell_coefficient or fouriernl or q1_poisson_golden_solution or tiled_executors_poisson" > gpurun_out/r02cg_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02cg_pytest.log; tail -n 25 gpurun_out/r02cg_pytest.log
