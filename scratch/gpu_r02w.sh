#!/bin/bash
# row-ordered vector executor: full GPU suite with it selected, then timings against the classic one
mkdir -p gpurun_out
T=${1:-r02w}
AFB_VEC_EXEC=rows timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -12 gpurun_out/${T}_pytest.log
for args in "140 1" "140 0" "203 1" "203 0"; do
  AFB_VEC_EXEC=rows timeout 300 python scratch/time_vec.py $args >> gpurun_out/${T}_time.log 2>&1
done
timeout 300 python scratch/time_vec.py 140 1 >> gpurun_out/${T}_time.log 2>&1
cat gpurun_out/${T}_time.log
