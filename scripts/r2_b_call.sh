#!/bin/bash
# Round 2, call b: per-layer bf16 parity harness on the GPU; bench lines of the other BASELINE configs (before the kernels
# are extended to S > 256 / L > 128: the "before" numbers); the eager (no graph) step.
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py -q -s 2>&1 | grep -v Warning | tail -60 > gpurun_out/r2_b_bf16_parity.log
tail -40 gpurun_out/r2_b_bf16_parity.log
for cfgs in "48 416" "128 448" "64 224" "64 320" "64 640"; do
  set -- $cfgs
  timeout 300 python bench.py --T $1 --res $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_bench_T$1_res$2.json 2> gpurun_out/r2_b_bench_T$1_res$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_b_bench_T$1_res$2.json")); e = d.get("encoder_attention", {})
    print("T$1 res$2", "ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "core us", e.get("us_core"), "block us", e.get("us_block"))
except Exception as ex:
    print("T$1 res$2 failed", ex)
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_b_bench_eager.json 2> gpurun_out/r2_b_bench_eager.err
cat gpurun_out/r2_b_bench_eager.json | cut -c1-600
