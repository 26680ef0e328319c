#!/bin/bash
# round-2 evidence: smoke, bench (both arms, driver defaults), launch list of the bench command, configs table
mkdir -p gpurun_out
T=${1:-r02z9}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench_n1.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2>> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e-pipeline --no-configs --no-first-step > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-200
timeout 900 python scratch/bench_configs.py c2 c3 c4 c5 > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err; tail -2 gpurun_out/${T}_configs.err
