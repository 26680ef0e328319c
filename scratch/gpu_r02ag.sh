#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02ag}
timeout 1500 python -m pytest tests -m gpu -x -q -k "tiled or poisson_values or elasticity or bilaplacian or degenerate or limits or full_size" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
python scratch/prof_inspector.py 120 1 > gpurun_out/${T}_wall.log 2>&1
python scratch/prof_inspector.py 203 3 >> gpurun_out/${T}_wall.log 2>&1
python scratch/prof_inspector.py 256 1 >> gpurun_out/${T}_wall.log 2>&1
cat gpurun_out/${T}_wall.log
timeout 600 python bench.py --n 120 --steps 10 --no-configs --no-cpu > gpurun_out/${T}_bench120.json 2> gpurun_out/${T}_bench120.err; tail -2 gpurun_out/${T}_bench120.err
python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench120.json').read().strip().splitlines()[-1]); print(d['phases']['add_and_compute_ms'], d['phases']['build_matrix_ms'], d['phases']['inspector_ms_once_per_mesh'], d['phases']['first_step_ms']['total_ms'], d['e2e']['new_mesh']['value'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_inspector_launches.csv python scratch/prof_inspector.py 120 1 > gpurun_out/${T}_ncu.log 2>&1
