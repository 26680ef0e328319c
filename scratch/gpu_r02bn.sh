#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02bn}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches_n120.csv python bench.py --n 120 --steps 2 --warmup 3 --no-cpu --no-e2e-pipeline --no-configs --no-first-step > gpurun_out/${T}_ncu_bench.log 2>&1
tail -c 300 gpurun_out/${T}_ncu_bench.log
timeout 600 python bench.py --n 128 --no-cpu --no-configs --no-first-step --no-e2e-pipeline > gpurun_out/${T}_bench_n128.json 2> gpurun_out/${T}_bench_n128.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_n128.json').read().strip().splitlines()[-1])
print('n128 value %.4e'%d['value'],'ms',round(d['ms_per_step'],4),'bm',round(d['phases']['build_matrix_ms'],4),'add',round(d['phases']['add_and_compute_ms'],4))
PY
