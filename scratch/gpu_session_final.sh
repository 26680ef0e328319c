#!/bin/bash
# round-end evidence: tests, bench (both arms), launch list of the bench command, full ncu of the top kernels, configs
mkdir -p gpurun_out
T=${1:-r01e}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench_n1.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2>> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_n120.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e-pipeline > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_assemble_tiled|k_pattern_nn_place|k_scan_chained" -s 6 -c 3 -o gpurun_out/${T}_full -f python scratch/prof_tiled.py > gpurun_out/${T}_ncu_full.log 2>&1; tail -1 gpurun_out/${T}_ncu_full.log
timeout 900 python scratch/bench_configs.py c2 c3 c4 c5 > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err; cut -c1-600 gpurun_out/${T}_configs.jsonl; tail -2 gpurun_out/${T}_configs.err
timeout 120 python scratch/prof_pcg.py 120 > gpurun_out/${T}_pcg_c2.json 2>&1; tail -1 gpurun_out/${T}_pcg_c2.json | cut -c1-300
