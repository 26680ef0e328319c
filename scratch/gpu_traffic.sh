#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02t}
for w in ${WORKLOADS:-c4 c3 c2 e3 e2 p2d}; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_traffic_$w.csv python scratch/prof_variants.py $w > gpurun_out/${T}_traffic_$w.log 2>&1
  tail -1 gpurun_out/${T}_traffic_$w.log | cut -c1-200; wc -l gpurun_out/${T}_traffic_$w.csv
done
