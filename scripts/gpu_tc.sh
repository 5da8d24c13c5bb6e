#!/bin/bash
# tcgen05 GEMM bring-up: kernel tests under a short timeout first (a protocol bug traps instead of hanging).
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q -s > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?" | tee -a gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
