#!/bin/bash
# final evidence on HEAD: full GPU suite (no -x), smoke, default bench + reference arm, ncu launch list + --set full captures
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_final_pytest_gpu.log | tail -4
bash scripts/gpu_profiles.sh 2>&1 | tail -25
