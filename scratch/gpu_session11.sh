#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/s11_smi.txt
for v in a57 a0 b0 b75 o0 a0b0 a0b0o0 a57b75; do
  echo "== $v" >> gpurun_out/s11_exp.log
  AFB200_LIB=$PWD/arcanefem_b200/variants/libafb200_$v.so timeout 120 python scratch/time_phases.py 120 2>&1 | grep tiled >> gpurun_out/s11_exp.log
done
echo "== base" >> gpurun_out/s11_exp.log
timeout 120 python scratch/time_phases.py 120 2>&1 | grep tiled >> gpurun_out/s11_exp.log
cat gpurun_out/s11_exp.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s11_pytest.log; tail -3 gpurun_out/s11_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; cat gpurun_out/s11_bench.json; tail -3 gpurun_out/s11_bench.err
