#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02f}
AFB_CHAIN_GEOM=F timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scratch/time_chain.py 48 > gpurun_out/${T}_memcheck.log 2>&1
grep -v "^\[W" gpurun_out/${T}_memcheck.log | head -60
