#!/bin/bash
# row-ordered vector executor: elasticity / tiled / multi-rank GPU tests with it selected, timings, one ncu capture
mkdir -p gpurun_out
T=${1:-r02y}
AFB_VEC_EXEC=rows timeout 1500 python -m pytest tests -m gpu -x -q -k "elasticity or tiled or degenerate or ownership or decomposed or mgpu or bilaplacian or facade" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
rm -f gpurun_out/${T}_time.log
for args in "140 1" "140 0" "203 1" "203 0"; do
  AFB_VEC_EXEC=rows timeout 300 python scratch/time_vec.py $args >> gpurun_out/${T}_time.log 2>&1
done
cat gpurun_out/${T}_time.log
AFB_VEC_EXEC=rows timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_rows_vec -s 2 -c 1 -o gpurun_out/${T}_rows -f python scratch/prof_vec.py 100 row > gpurun_out/${T}_rows.log 2>&1
tail -1 gpurun_out/${T}_rows.log
