#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
for prio in 1 0; do
  STCAT_CHAIN_PRIO=$prio timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_r_bench_prio$prio.json 2> gpurun_out/r2_r_bench_prio$prio.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_r_bench_prio$prio.json"))
    print("prio $prio: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as ex:
    print("prio $prio failed", ex); print(open("gpurun_out/r2_r_bench_prio$prio.err").read()[-800:])
PY
done
