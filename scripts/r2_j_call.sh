#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_hotpath.py -q -W ignore -x 2>&1 | tail -4
for sms in 0 128 112 96 80; do
  STCAT_MEMSIDE_SMS=$sms timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_j_bench_sms$sms.json 2> gpurun_out/r2_j_bench_sms$sms.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_j_bench_sms$sms.json")); e = d.get("encoder_attention", {})
    print("memside sms $sms: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "block us", round(e.get("us_block", 0), 1), "roofline frac", round(d["roofline"]["frac"], 3), "us", round(d["roofline"]["us_per_launch"], 1))
except Exception as ex:
    print("sms $sms failed", ex)
PY
done
STCAT_GEMM_EPI2=0 STCAT_MEMSIDE_SMS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_j_bench_noepi2.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/r2_j_bench_noepi2.json"))
print("EPI2 off, no cap: ms", round(d["ms_per_step"], 3))
PY
timeout 120 python scripts/bench_gemm.py 2>/dev/null | tee gpurun_out/r2_j_gemm_table.txt | head -30
