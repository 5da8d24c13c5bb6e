#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -W ignore -x 2>&1 | tail -2
timeout 200 python scripts/bench_gemm.py 2>/dev/null | tee gpurun_out/r2_o_gemm_table.txt | tail -28
echo "== BN=64 forced"
for n in out_fwd v_fwd out_dgrad qk_fwd mem_kv_fwd; do STCAT_TC_BN=64 timeout 100 python scripts/bench_gemm.py --only $n 2>/dev/null | tail -1; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_o_bench.json 2> gpurun_out/r2_o_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_o_bench.json")); e = d.get("encoder_attention", {})
print("ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "block us", round(e.get("us_block", 0), 1), "family", d["gemm_family"].get("frac_of_bf16_peak"))
PY
