#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log; tail -2 gpurun_out/full_pytest.log
timeout 300 python scratch/bench_configs.py c3 2>&1 | tee gpurun_out/c3_final.jsonl | python -c "
import sys, json
for l in sys.stdin:
    c = json.loads(l); print(c['config'], {k: (round(v['build_matrix_ms'], 3), round(v['add_and_compute_ms'], 3), round(v['values_frac_of_peak'], 3)) for k, v in c['variants'].items()})
"
