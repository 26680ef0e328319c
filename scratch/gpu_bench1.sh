#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02n}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "update_coordinates" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d['e2e'])[:1500])
print(d['check']['ok'], [ (c['config'], round(c['roofline']['frac'],3), c['check']['ok']) for c in d['configs']])
PY
tail -3 gpurun_out/${T}_bench.err
