#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu -k "single_query" > gpurun_out/r2_y_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_y_pytest.log | tail -8
timeout 120 python scripts/bench_attn_sq.py 2>&1 | tee gpurun_out/r2_y_attn_sq.txt
