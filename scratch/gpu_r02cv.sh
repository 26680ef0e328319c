#!/bin/bash
# final build of the round: smoke + full GPU suite + north-star bench (C4 default) + reference arm
mkdir -p gpurun_out
T=r02cv
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -n 4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1_c4.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench_n1_c4.json; tail -n 3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2>> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_reference_arm.json
