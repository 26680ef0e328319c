#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02av}
rm -f gpurun_out/${T}_time.log
for args in "140 0" "203 0" "203 1"; do
  timeout 300 python scratch/time_vec.py $args >> gpurun_out/${T}_time.log 2>&1
done
cat gpurun_out/${T}_time.log
timeout 900 python -m pytest tests -m gpu -x -q -k "elasticity or full_size or mgpu or cut_rows or decomposed" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
