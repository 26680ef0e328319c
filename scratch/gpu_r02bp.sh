#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02bp}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
for n in 120 256; do for env in "AFB_NO_PDL=0" "AFB_NO_PDL=1"; do
  env $env timeout 900 python bench.py --n $n --no-cpu --no-first-step --no-e2e-pipeline > gpurun_out/${T}_bench_n${n}_${env}.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_n*json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], 'value %.4e'%d['value'],'ms',round(d['ms_per_step'],4),'bm',round(d['phases']['build_matrix_ms'],4),'add',round(d['phases']['add_and_compute_ms'],4),'launches',d['gpu_launches'], d['check']['ok'], [(c.get('name','?')[:10], round(c.get('ms',0),4)) for c in d.get('configs',[])] if isinstance(d.get('configs'),list) else list(d.get('configs',{}).keys()))
PY
