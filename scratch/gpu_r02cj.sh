#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02cj_traffic_coef_c2.csv python scratch/time_coef.py 120 > gpurun_out/r02cj_traffic_coef_c2.log 2>&1
tail -1 gpurun_out/r02cj_traffic_coef_c2.log | cut -c1-300; wc -l gpurun_out/r02cj_traffic_coef_c2.csv
