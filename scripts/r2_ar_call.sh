#!/bin/bash
# box head + refinement + sine as one launch, query_sine written as operand: kernel tests, parity, train-step tests, bench + trace
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_hotpath.py tests/test_gpu_bf16_parity.py tests/test_gpu_train_step.py -q -m gpu -k "box_head or fused_glue or every_layer or train or replay or leaf or anchor" > gpurun_out/r2_ar_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_ar_pytest.log | tail -8
run() {
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_ar_profile_$tag.md > gpurun_out/r2_ar_bench_$tag.json 2> gpurun_out/r2_ar_bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_ar_bench_$tag.json"))
    print("$tag step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"], "loss", d["e2e"].get("loss"))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_ar_bench_$tag.err").read()[-1500:])
PY
}
run head STCAT_TRACE=gpurun_out/r2_ar_trace.json
gzip -f gpurun_out/r2_ar_trace.json
