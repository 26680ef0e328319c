#!/bin/bash
for L in 2 3 4; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-lanes $L > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['one_step_at_a_time_value'], d['e2e']['pipelined'][:10])
PY
tail -2 gpurun_out/q_bench.err
done
