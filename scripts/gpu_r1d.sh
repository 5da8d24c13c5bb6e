#!/bin/bash
# Session check: parity tests, smoke, bench + trace, GEMM micro-bench, ncu --set full on the FFN1 GEMM.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_gemm.py > gpurun_out/gemm_table.txt 2>&1; echo "gemm rc=$?"
cat gpurun_out/gemm_table.txt
STCAT_TRACE=gpurun_out/trace.json timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline --profile gpurun_out/profile_bf16.md > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_bf16.json; grep -v Warning gpurun_out/bench_bf16.err | tail -5
gzip -f gpurun_out/trace.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 2 -f -o gpurun_out/prof_ffn1 \
    python scripts/bench_gemm.py --only ffn1_fwd --iters 2 > gpurun_out/ncu_ffn1.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_ffn1.log
ls -la gpurun_out
