#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "heat or msh_driver or native_msh" > gpurun_out/r02cd_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02cd_pytest.log; tail -n 5 gpurun_out/r02cd_pytest.log
