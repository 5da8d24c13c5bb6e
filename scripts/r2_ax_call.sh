#!/bin/bash
# bench lines of every BASELINE configuration on HEAD (end of round 2)
mkdir -p gpurun_out
export PYTHONPATH=.
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
run() { name=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_ax_bench_$name.json 2> gpurun_out/r2_ax_bench_$name.err; show gpurun_out/r2_ax_bench_$name.json "$name"; }
run T48_res416 --T 48 --res 416
run T128_res448 --T 128
run T64_res224 --res 224
run T64_res320 --res 320
run T64_res640 --res 640
run T200_res448 --T 200
run T64_res448_dropout01 --dropout 0.1
