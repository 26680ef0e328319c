#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cell_coefficient or fouriernl" > gpurun_out/r02ci_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ci_pytest.log; tail -n 5 gpurun_out/r02ci_pytest.log
timeout 300 python scratch/time_coef.py 120 256 > gpurun_out/r02ci_time_coef.jsonl 2> gpurun_out/r02ci_time_coef.err; cat gpurun_out/r02ci_time_coef.jsonl; tail -3 gpurun_out/r02ci_time_coef.err
