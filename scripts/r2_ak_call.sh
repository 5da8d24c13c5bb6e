#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
STCAT_NO_PDL=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dropout 0.1 --profile gpurun_out/r2_ak_profile_dropout_nopdl.md > gpurun_out/r2_ak_bench.json 2> gpurun_out/r2_ak_bench.err
grep "attn_tc\|drop_bits\|dropout_kernel\|attn_mma\|attn_sq\|layernorm\|# 3 steps" gpurun_out/r2_ak_profile_dropout_nopdl.md | cut -c1-160
