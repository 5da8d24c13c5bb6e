#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_hotpath.py -q -W ignore -x 2>&1 | tail -6
for rows in 0 1024 100000; do
  STCAT_LEAF_ROWS=$rows timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_p_bench_leaf$rows.json 2> gpurun_out/r2_p_bench_leaf$rows.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_p_bench_leaf$rows.json"))
    print("leaf rows $rows: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "loss", d["e2e"]["loss"])
except Exception as ex:
    print("leaf $rows failed", ex); print(open("gpurun_out/r2_p_bench_leaf$rows.err").read()[-800:])
PY
done
