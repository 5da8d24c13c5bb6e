#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite on HEAD without -x, then the staged variants.
mkdir -p gpurun_out
export PYTHONPATH=.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_a_pytest_gpu.log
tail -15 gpurun_out/r2_a_pytest_gpu.log
timeout 900 bash scripts/validate_staged.sh > gpurun_out/r2_a_validate_staged.log 2>&1
cat gpurun_out/r2_a_validate_staged.log
