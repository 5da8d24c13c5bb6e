#!/bin/bash
# bbox_embed + refinement + sine as one autograd node (ops.box_mlp_head): kernel tests, parity, train-step tests, A/B on one box
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_hotpath.py tests/test_gpu_bf16_parity.py tests/test_gpu_train_step.py -q -m gpu -k "box_head or fused_glue or every_layer or train or replay or leaf or anchor" > gpurun_out/r2_at_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_at_pytest.log | tail -6
run() {
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/r2_at_bench_$tag.json 2> gpurun_out/r2_at_bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_at_bench_$tag.json"))
    print("$tag step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"], "loss", d["e2e"].get("loss"))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_at_bench_$tag.err").read()[-1500:])
PY
}
run head1 STCAT_FUSED_HEAD=1
run head0 STCAT_FUSED_HEAD=0
run head1b STCAT_FUSED_HEAD=1
run unfused STCAT_FUSED_GLUE=0 STCAT_ZERO_ASYNC=0
