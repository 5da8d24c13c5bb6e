#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
STCAT_TRACE=gpurun_out/r2_ac_trace.json.gz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_ac_profile.md > gpurun_out/r2_ac_bench.json 2> gpurun_out/r2_ac_bench.err
python scripts/trace_timeline.py gpurun_out/r2_ac_trace.json.gz 250 | head -30
STCAT_NO_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile gpurun_out/r2_ac_profile_nopdl.md > gpurun_out/r2_ac_bench_nopdl.json 2> gpurun_out/r2_ac_bench_nopdl.err
head -3 gpurun_out/r2_ac_profile_nopdl.md
timeout 2400 bash scripts/sanitize.sh memcheck 2>&1 | tail -12
