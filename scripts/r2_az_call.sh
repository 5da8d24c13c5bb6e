#!/bin/bash
# frame-CLS exchange kernels with fused operand copies + early q/k/v projection under the temporal layer: tests, A/B with phases
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_hotpath.py tests/test_gpu_bf16_parity.py tests/test_gpu_train_step.py -q -m gpu -k "cls_gather or fused_glue or every_layer or train or replay or leaf or fast_kernels" > gpurun_out/r2_az_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_az_pytest.log | tail -5
run() {
tag=$1; shift
env "$@" timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --phases > gpurun_out/r2_az_bench_$tag.json 2> gpurun_out/r2_az_bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_az_bench_$tag.json"))
    print("$tag step: ms", round(d["ms_per_step"], 3), "launches/step", d["gpu_launches"] / d["steps"], "loss", d["e2e"].get("loss"))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_az_bench_$tag.err").read()[-1500:])
PY
grep phases gpurun_out/r2_az_bench_$tag.err
}
run default
run early0 STCAT_EARLY_QKV=0
run early0_cls0 STCAT_EARLY_QKV=0 STCAT_CLS_KERNELS=0
