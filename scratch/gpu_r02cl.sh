#!/bin/bash
mkdir -p gpurun_out
T=r02cl
timeout 600 python -m pytest tests -m gpu -x -q -k "cell_coefficient or fouriernl or q1_poisson_golden_solution or poisson_values or elasticity_values" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -n 3 gpurun_out/${T}_pytest.log
timeout 300 python scratch/time_coef.py 120 256 > gpurun_out/${T}_time_coef.jsonl 2> gpurun_out/${T}_time_coef.err; cat gpurun_out/${T}_time_coef.jsonl; tail -3 gpurun_out/${T}_time_coef.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1_c4.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench_n1_c4.json; tail -n 3 gpurun_out/${T}_bench.err
