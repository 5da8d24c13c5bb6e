#!/bin/bash
# full GPU suite on HEAD (no -x), then the other BASELINE configurations, dropout and eager steps
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_ae_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_ae_pytest_gpu.log | tail -6
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    ea = d.get("encoder_attention", {})
    print(sys.argv[2], ": ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "attn core us", round(ea.get("us_core", 0), 1), "block us", round(ea.get("us_block", 0), 1))
except Exception as ex:
    print(sys.argv[2], "failed", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
run() { name=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_ae_bench_$name.json 2> gpurun_out/r2_ae_bench_$name.err; show gpurun_out/r2_ae_bench_$name.json "$name"; }
run T64_res448
run T48_res416 --T 48 --res 416
run T128_res448 --T 128
run T64_res224 --res 224
run T64_res320 --res 320
run T64_res640 --res 640
run dropout01 --dropout 0.1
run eager --no-graph
