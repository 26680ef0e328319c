#!/bin/bash
mkdir -p gpurun_out
T=r02cx
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1_c4.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench_n1_c4.json; tail -n 3 gpurun_out/${T}_bench.err
