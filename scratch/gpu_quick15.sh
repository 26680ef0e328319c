#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "poisson_values or golden_solution or full_size_poisson or is_own or ownership or degenerate or signed_area or decomposed" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -5 gpurun_out/q_pytest.log
timeout 300 python scratch/time_phases.py 120 2>&1 | grep -E "atomic" | cut -c1-140
timeout 300 python scratch/bench_configs.py c5 2>&1 | grep -E "P1 b=1" | python -c "
import sys, json
for l in sys.stdin:
    c = json.loads(l); print(c['config'], {k: round(v['add_and_compute_ms'], 3) for k, v in c['variants'].items()})
"
