#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02d}
for cfg in "chain F 1" "chain F 0"; do
  set -- $cfg
  for n in 24 120 256; do
    AFB_SCALAR_EXEC=$1 AFB_CHAIN_GEOM=$2 AFB_CHAIN_PREFILL=$3 timeout 120 python scratch/time_chain.py $n 2>&1 | tail -1 | cut -c1-600
  done
done > gpurun_out/${T}_time.log 2>&1
cat gpurun_out/${T}_time.log
AFB_CHAIN_GEOM=F timeout 120 python scratch/time_chain.py 2048 2 2>&1 | tail -1 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -15 gpurun_out/${T}_pytest.log
