#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log; tail -12 gpurun_out/t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
