#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/time_coef.py 120 256 > gpurun_out/r02ch_time_coef.jsonl 2> gpurun_out/r02ch_time_coef.err; cat gpurun_out/r02ch_time_coef.jsonl; tail -3 gpurun_out/r02ch_time_coef.err
