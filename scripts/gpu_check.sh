#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both precisions), ncu launch list.  Usage (here):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 --cpu-steps 1 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?"
tail -c 3000 gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?"
tail -c 3000 gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --precision ${NCU_PRECISION:-bf16} > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
