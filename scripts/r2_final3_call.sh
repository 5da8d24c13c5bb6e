#!/bin/bash
# last check on HEAD: full GPU suite (no -x) and the default-size bench line
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r2_final3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_final3_pytest_gpu.log | tail -3
timeout 60 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final3_bench_1gpu.json 2> gpurun_out/r2_final3_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final3_bench_1gpu.json"))
print("HEAD: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"], d["clocks"])
PY
