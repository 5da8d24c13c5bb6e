#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_hotpath.py tests/test_gpu_train_step.py -q -m gpu > gpurun_out/r2_ad_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_ad_pytest.log | tail -6
STCAT_TRACE=gpurun_out/r2_ad_trace.json.gz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_ad_profile.md > gpurun_out/r2_ad_bench_tr.json 2> gpurun_out/r2_ad_bench_tr.err
python scripts/trace_timeline.py gpurun_out/r2_ad_trace.json.gz 250 | head -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_ad_bench.json 2> gpurun_out/r2_ad_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_ad_bench.json"))
    print("step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"])
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_ad_bench.err").read()[-1500:])
PY
timeout 1200 bash scripts/sanitize.sh racecheck synccheck 2>&1 | tail -16
