#!/bin/bash
# memory-side projections of a box + time decoder layer pair as one grouped launch: parity, train-step tests, A/B on one box
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_bf16_parity.py tests/test_gpu_train_step.py -q -m gpu -k "fused_glue or every_layer or train or replay or leaf or grad_fusion or golden" > gpurun_out/r2_au_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_au_pytest.log | tail -6
run() {
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_au_profile_$tag.md > gpurun_out/r2_au_bench_$tag.json 2> gpurun_out/r2_au_bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_au_bench_$tag.json"))
    print("$tag step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"], "loss", d["e2e"].get("loss"))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_au_bench_$tag.err").read()[-1500:])
PY
}
run group1 STCAT_GROUP_MEMSIDE=1 STCAT_TRACE=gpurun_out/r2_au_trace.json
run group0 STCAT_GROUP_MEMSIDE=0
run group1_nocap STCAT_GROUP_MEMSIDE=1 STCAT_MEMSIDE_SMS=0
gzip -f gpurun_out/r2_au_trace.json
