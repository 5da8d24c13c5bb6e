#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1200 python -m pytest tests/test_gpu_bf16_parity.py -q -s -W ignore 2>&1 > gpurun_out/r2_c_bf16_parity.log
grep -n "^\[\|passed\|failed\|FAILED\|Error" gpurun_out/r2_c_bf16_parity.log | cut -c1-400 | tail -60
