#!/bin/bash
# first GPU session of this container: tests, bench, phase timing, launch list, full ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 120 python scratch/phase_tiled.py > gpurun_out/phase.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tiled.csv python bench.py --steps 2 --warmup 3 --no-cpu --variant tiled > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiled -s 2 -c 1 -o gpurun_out/prof_tiled python scratch/prof_tiled.py > gpurun_out/ncu_tiled.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pattern_rows -s 1 -c 1 -o gpurun_out/prof_pattern python scratch/prof_pattern.py > gpurun_out/ncu_pattern.log 2>&1
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench.json; cat gpurun_out/phase.log
