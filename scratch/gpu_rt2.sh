#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "elasticity or bilaplacian or tiled_pattern" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -2 gpurun_out/q_pytest.log
timeout 600 python scratch/bench_configs.py c5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    c = json.loads(l); print(c['config'], {k: (round(v['build_matrix_ms'], 3), round(v['add_and_compute_ms'], 3)) for k, v in c['variants'].items()})
"
