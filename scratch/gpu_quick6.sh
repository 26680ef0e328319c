#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q -k "peer_memory" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -30 gpurun_out/q_pytest.log
