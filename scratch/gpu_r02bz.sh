#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02bz}
python scratch/prof_inspector3.py 120 120 256 120 2> gpurun_out/${T}_inspector_trace.txt; grep "first tiled" gpurun_out/${T}_inspector_trace.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n1_c4.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_n1_c4.json').read().strip().splitlines()[-1])
print('value %.4e'%d['value'],'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'e2e %.3e'%d['e2e']['value'], d['phases'].get('first_step_ms'), d['phases'].get('inspector_ms_once_per_mesh'), 'cpu %.3e'%d['cpu_baseline']['value'])
PY
