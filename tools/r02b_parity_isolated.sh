#!/bin/bash
# tests/test_gpu_parity.py under pytest-xdist with one worker: a crashing test takes only the worker down and is named
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 500 python -m pytest tests/test_gpu_parity.py -q -n 1 --max-worker-restart 4 > gpurun_out/v_parity_full.txt 2>&1
echo "rc=$?"
grep -n "crashed\|FAILED\|passed\|failed\|Error" gpurun_out/v_parity_full.txt | head -20
