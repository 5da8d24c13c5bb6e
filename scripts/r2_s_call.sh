#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
for sms in 0 112 80 48; do
  STCAT_LEAF_SMS=$sms timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_s_bench_$sms.json 2> gpurun_out/r2_s_bench_$sms.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_s_bench_$sms.json"))
    print("leaf sms $sms: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1))
except Exception as ex:
    print("$sms failed", ex); print(open("gpurun_out/r2_s_bench_$sms.err").read()[-800:])
PY
done
