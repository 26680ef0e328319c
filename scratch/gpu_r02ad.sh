#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02ad}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python scratch/bench_configs.py c3 c5 > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err; tail -3 gpurun_out/${T}_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/%s_configs.jsonl' % "${T}"):
    try: d=json.loads(l)
    except Exception: continue
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','variant','layout','build_matrix_ms','add_and_compute_ms','frac','gbs')})
PY
