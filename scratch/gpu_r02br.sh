#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02br}
for n in 120 256; do for lib in libafb200.so variants/libafb200_stcs.so; do
  AFB200_LIB=$PWD/arcanefem_b200/$lib timeout 900 python bench.py --n $n --no-cpu --no-configs --no-first-step --no-e2e-pipeline > gpurun_out/${T}_bench_n${n}_$(basename $lib .so).json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_n*json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], 'value %.4e'%d['value'],'ms',round(d['ms_per_step'],4),'bm',round(d['phases']['build_matrix_ms'],4),'add',round(d['phases']['add_and_compute_ms'],4), d['check']['ok'])
PY
