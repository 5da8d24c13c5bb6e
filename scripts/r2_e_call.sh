#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_attention_tc.py -q -s -W ignore -x 2>&1 > gpurun_out/r2_e_attn_tc.log
grep -n "B=\|passed\|failed\|FAILED\|Error\|us/launch" gpurun_out/r2_e_attn_tc.log | cut -c1-300 | tail -50
timeout 120 python scripts/attn_timeline.py 64 213 fwd > gpurun_out/r2_e_attn_fwd_timeline.txt 2>&1; tail -9 gpurun_out/r2_e_attn_fwd_timeline.txt
timeout 120 python scripts/attn_timeline.py 64 213 bwd > gpurun_out/r2_e_attn_bwd_timeline.txt 2>&1; tail -36 gpurun_out/r2_e_attn_bwd_timeline.txt
