#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -5
timeout 300 python -m pytest tests/test_gpu_optim.py -q -W ignore 2>&1 | tail -2
STCAT_TRACE=gpurun_out/r2_h_trace.json.gz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_h_profile.md > gpurun_out/r2_h_bench.json 2> gpurun_out/r2_h_bench.err
tail -3 gpurun_out/r2_h_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_h_bench.json")); e = d.get("encoder_attention", {})
print("with optimizer: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"] / d["steps"], "core us", e.get("us_core"), "block us", e.get("us_block"))
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer > gpurun_out/r2_h_bench_noopt.json 2> gpurun_out/r2_h_bench_noopt.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_h_bench_noopt.json"))
print("no optimizer: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1))
PY
