#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_kernels.py tests/test_gpu_attention_tc.py -q -W ignore -x 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_i_bench.json 2> gpurun_out/r2_i_bench.err; tail -2 gpurun_out/r2_i_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dropout 0.1 > gpurun_out/r2_i_bench_dropout01.json 2> gpurun_out/r2_i_bench_dropout01.err; tail -2 gpurun_out/r2_i_bench_dropout01.err
python - <<PY
import json
for f in ("r2_i_bench", "r2_i_bench_dropout01"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "graph", d["config"]["cuda_graph"])
    except Exception as e:
        print(f, "failed", e)
PY
