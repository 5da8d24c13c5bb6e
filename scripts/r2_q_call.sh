#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
STCAT_TRACE=gpurun_out/r2_q_trace.json.gz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_q_profile.md > gpurun_out/r2_q_bench.json 2> gpurun_out/r2_q_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_q_bench.json"))
print("ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
PY
python scripts/trace_timeline.py gpurun_out/r2_q_trace.json.gz 250 | head -30
