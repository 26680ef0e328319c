#!/bin/bash
# full GPU test-suite, smoke, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log; tail -3 gpurun_out/full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; cut -c1-330 gpurun_out/full_bench.json; tail -2 gpurun_out/full_bench.err
