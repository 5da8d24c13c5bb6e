#!/bin/bash
# all GPU parity tests, then bench + profile + trace
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/pytest_gpu.log | tail -${PYTAIL:-8}
TESTS= bash scripts/gpu_bench.sh
