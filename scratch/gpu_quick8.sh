#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "tiled or elasticity or ownership or decomposed" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log
timeout 300 python scratch/time_phases.py 120 2>&1 | grep conn | cut -c1-120
timeout 120 python scratch/prof_vec.py 100 2>&1 | tail -1 | cut -c1-120
timeout 120 python scratch/prof_vec.py 100 row 2>&1 | tail -1 | cut -c1-120
