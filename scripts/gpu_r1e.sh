#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -5 gpurun_out/pytest_tc.log
timeout 300 python scripts/bench_gemm.py > gpurun_out/gemm_table.txt 2>&1; echo "gemm rc=$?"
cat gpurun_out/gemm_table.txt
timeout 300 python scripts/probe_determinism.py > gpurun_out/determinism.txt 2>&1; grep -v Warn gpurun_out/determinism.txt | tail -8
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
