#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention_tc.py -x -q -s > gpurun_out/pytest_attn.log 2>&1; echo "attn rc=$?" | tee -a gpurun_out/pytest_attn.log
tail -40 gpurun_out/pytest_attn.log
