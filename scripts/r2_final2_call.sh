#!/bin/bash
# final evidence on HEAD (second half of round 2): full GPU suite (no -x), smoke, default bench + reference arm, ncu launch list +
# --set full captures (scripts/gpu_profiles.sh), then the bench lines of every BASELINE configuration
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_final2_pytest_gpu.log | tail -4
bash scripts/gpu_profiles.sh 2>&1 | tail -25
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --phases > gpurun_out/r2_final2_bench_phases.json 2> gpurun_out/r2_final2_bench_phases.err; grep phases gpurun_out/r2_final2_bench_phases.err
bash scripts/r2_ax_call.sh
