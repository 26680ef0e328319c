#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "elasticity or tiled or ownership or distributed or emulated" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -4 gpurun_out/q_pytest.log
timeout 120 python scratch/prof_vec.py 100 2>&1 | tail -1
timeout 120 python scratch/prof_vec.py 100 row 2>&1 | tail -1
