#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02ao}
timeout 900 python -m pytest tests/test_cpp_facade.py tests/test_distributed.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
bash scratch/gpu_multi2.sh 4 $T > /dev/null; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases']['per_rank_ms'], d['e2e']['value'], d['check']['ok'])"
tail -3 gpurun_out/${T}_bench_n4.err
