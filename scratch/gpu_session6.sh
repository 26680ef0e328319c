#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
echo "== default" > gpurun_out/phases.log; timeout 300 python scratch/time_phases.py 120 100 >> gpurun_out/phases.log 2>&1
for v in B C D F; do
  echo "== variant $v" >> gpurun_out/phases.log
  AFB200_LIB=$PWD/arcanefem_b200/variants/libafb200_$v.so timeout 300 python scratch/time_phases.py 120 100 >> gpurun_out/phases.log 2>&1
  AFB200_LIB=$PWD/arcanefem_b200/variants/libafb200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "poisson_values or elasticity_values or tiled or ownership" > gpurun_out/pytest_$v.log 2>&1; echo "variant $v pytest: $(tail -1 gpurun_out/pytest_$v.log)" >> gpurun_out/phases.log
done
cat gpurun_out/phases.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiled -s 2 -c 1 -o gpurun_out/prof_tiled python scratch/prof_tiled.py > gpurun_out/ncu_tiled.log 2>&1
