#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pattern" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log
timeout 300 python scratch/time_phases.py 120 2>&1 | grep conn
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print(d['value'], d['ms_per_step'], d['phases']['build_matrix_ms'], d['phases']['add_and_compute_ms'], d['e2e'])
PY
tail -3 gpurun_out/q_bench.err
