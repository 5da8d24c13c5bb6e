#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_hotpath.py tests/test_gpu_kernels.py -q -W ignore -x 2>&1 | tail -4
timeout 200 python scripts/bench_gemm.py 2>/dev/null | tee gpurun_out/r2_m_gemm_table.txt | tail -32
STCAT_GEMM_DIRECT=0 timeout 200 python scripts/bench_gemm.py 2>/dev/null | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_m_bench.json 2> gpurun_out/r2_m_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_m_bench.json")); e = d.get("encoder_attention", {})
print("ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "block us", round(e.get("us_block", 0), 1), "core", round(e.get("us_core", 0), 1), "roofline frac", round(d["roofline"]["frac"], 3), "us", round(d["roofline"]["us_per_launch"], 1), "family", d["gemm_family"].get("frac_of_bf16_peak"))
PY
