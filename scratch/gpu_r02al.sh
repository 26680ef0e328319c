#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02al}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
print('insp', d['phases']['inspector_ms_once_per_mesh'], 'first', d['phases']['first_step_ms'])
print('e2e', d['e2e']['value'], d['e2e']['new_mesh']['value'])
for c in d['configs']: print(c['config'], c['add_and_compute_ms'], c['roofline']['frac'], c['inspector_ms_once_per_mesh'])
PY
