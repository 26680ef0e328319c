#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 900 python scratch/bench_configs.py c2 c3 c4 c5 > gpurun_out/configs.log 2>&1; cat gpurun_out/configs.log | cut -c1-1500
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python scratch/prof_tiled.py > gpurun_out/ncu_l.log 2>&1
