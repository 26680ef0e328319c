#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02b}
for cfg in "chain B 1" "chain A 1" "chain B 0" "tiles B 1"; do
  set -- $cfg
  for n in 24 120 256; do
    AFB_SCALAR_EXEC=$1 AFB_CHAIN_GEOM=$2 AFB_CHAIN_PREFILL=$3 timeout 300 python scratch/time_chain.py $n 2>&1 | tail -1 | cut -c1-600
  done
done > gpurun_out/${T}_time.log 2>&1
cat gpurun_out/${T}_time.log
AFB_CHAIN_GEOM=B timeout 300 python scratch/time_chain.py 2048 2 2>&1 | tail -1 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -15 gpurun_out/${T}_pytest.log
