#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_bf16_parity.py -q -s -m gpu 2>&1 | grep "per-layer forward\|passed\|failed\|end to end" | tee gpurun_out/r2_aj_bf16_parity.log
