#!/bin/bash
mkdir -p gpurun_out
AFB_TILED_EXEC=pipe timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_distributed.py -m gpu -x -q -k "tiled or poisson_values or ownership or full_size_poisson or decomposed or golden" > gpurun_out/pytest_pipe.log 2>&1; echo "pipe pytest rc=$?" >> gpurun_out/pytest_pipe.log; tail -5 gpurun_out/pytest_pipe.log
echo "== phase" > gpurun_out/phases.log; AFB_TILED_EXEC=phase timeout 200 python scratch/time_phases.py 120 >> gpurun_out/phases.log 2>&1
echo "== pipe" >> gpurun_out/phases.log; AFB_TILED_EXEC=pipe timeout 200 python scratch/time_phases.py 120 >> gpurun_out/phases.log 2>&1
cut -c1-100 gpurun_out/phases.log
AFB_TILED_EXEC=pipe timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiled -s 2 -c 1 -o gpurun_out/prof_pipe python scratch/prof_tiled.py > gpurun_out/ncu_pipe.log 2>&1
AFB_TILED_EXEC=phase timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiled -s 2 -c 1 -o gpurun_out/prof_tiled python scratch/prof_tiled.py > gpurun_out/ncu_tiled.log 2>&1
