#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpp_facade.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -12 gpurun_out/q_pytest.log
