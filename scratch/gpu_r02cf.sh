#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fouriernl" > gpurun_out/r02cf_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02cf_pytest.log; tail -n 25 gpurun_out/r02cf_pytest.log
