#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pcg" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log
python scratch/prof_pcg.py 120 | tail -1
