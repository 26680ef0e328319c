#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "aerodynamics" > gpurun_out/r02ce_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ce_pytest.log; tail -n 15 gpurun_out/r02ce_pytest.log
