#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1500 python -m pytest tests -q -m gpu -W ignore 2>&1 | tail -15 > gpurun_out/r2_f_pytest_gpu.log
tail -8 gpurun_out/r2_f_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_f_bench.json 2> gpurun_out/r2_f_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_f_bench.json")); e = d.get("encoder_attention", {})
print("ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", d["e2e"]["value"], "core us", e.get("us_core"), "block us", e.get("us_block"), "roofline", d["roofline"])
PY
for cfgs in "64 640" "128 448"; do
  set -- $cfgs
  timeout 300 python bench.py --T $1 --res $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_f_bench_T$1_res$2.json 2> gpurun_out/r2_f_bench_T$1_res$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_f_bench_T$1_res$2.json")); e = d.get("encoder_attention", {})
    print("T$1 res$2", "ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "core us", e.get("us_core"), "block us", e.get("us_block"))
except Exception as ex:
    print("T$1 res$2 failed", ex)
PY
done
