#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "q1_elasticity" > gpurun_out/r02cy_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02cy_pytest.log; tail -n 30 gpurun_out/r02cy_pytest.log
