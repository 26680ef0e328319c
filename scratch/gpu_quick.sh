#!/bin/bash
# quick loop: tiled parity subset + phase timings (+ optional extra args: sizes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "tiled or Tiled or pattern" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -4 gpurun_out/q_pytest.log
timeout 300 python scratch/time_phases.py 120 ${1:-} 2>&1 | tee gpurun_out/q_time.log
