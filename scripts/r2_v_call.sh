#!/bin/bash
# CLS-row split of the spatial encoder layers: parity tests, then the step with / without it
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_hotpath.py tests/test_gpu_train_step.py -q -m gpu > gpurun_out/r2_v_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_v_pytest.log | tail -15
for ov in 1 0; do
  STCAT_CLS_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_v_bench_ov$ov.json 2> gpurun_out/r2_v_bench_ov$ov.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_v_bench_ov$ov.json"))
    print("cls overlap $ov: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as ex:
    print("$ov failed", ex); print(open("gpurun_out/r2_v_bench_ov$ov.err").read()[-1500:])
PY
done
STCAT_TRACE=gpurun_out/r2_v_trace.json.gz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_v_profile.md > gpurun_out/r2_v_bench_trace.json 2> gpurun_out/r2_v_bench_trace.err
python scripts/trace_timeline.py gpurun_out/r2_v_trace.json.gz 250 | head -30
