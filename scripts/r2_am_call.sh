#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_hotpath.py tests/test_gpu_train_step.py -q -m gpu > gpurun_out/r2_am_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_am_pytest.log | tail -4
for dp in 0 0.1; do
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dropout $dp > gpurun_out/r2_am_bench_$dp.json 2> gpurun_out/r2_am_bench_$dp.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_am_bench_$dp.json"))
    print("dropout $dp step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"])
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_am_bench_$dp.err").read()[-1500:])
PY
done
