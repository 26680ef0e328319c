#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02bl}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
for env in "" "AFB_NO_FUSED_COLUMNS=1"; do
  env $env timeout 900 python bench.py --no-cpu --no-configs --no-first-step > gpurun_out/${T}_bench_${env:-fused}.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
  env $env timeout 900 python bench.py --no-cpu --no-configs --no-first-step --n 120 > gpurun_out/${T}_bench120_${env:-fused}.json 2>> gpurun_out/${T}_bench.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench*json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], 'value %.4e'%d['value'],'ms',round(d['ms_per_step'],4),'bm',round(d['phases']['build_matrix_ms'],4),'add',round(d['phases']['add_and_compute_ms'],4),'frac',round(d['roofline']['frac'],4), d['phases'].get('columns_written_by_the_assembly_kernel'), d['check']['ok'], 'e2e %.3e'%d['e2e']['value'])
PY
