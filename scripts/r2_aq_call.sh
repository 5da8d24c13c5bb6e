#!/bin/bash
# gradient sink of the decoder's memory operands (ops._GradSink): GEMM bf16-accumulate tests, model-level tests, bench A/B + trace
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_hotpath.py tests/test_gpu_bf16_parity.py tests/test_gpu_train_step.py -q -m gpu -k "bwd_bf16 or fused_glue or every_layer or train or replay or leaf" > gpurun_out/r2_aq_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_aq_pytest.log | tail -8
run() {
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2_aq_profile_$tag.md > gpurun_out/r2_aq_bench_$tag.json 2> gpurun_out/r2_aq_bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_aq_bench_$tag.json"))
    print("$tag step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"], "loss", d["e2e"].get("loss"))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_aq_bench_$tag.err").read()[-1500:])
PY
}
run sink STCAT_GRAD_SINK=1 STCAT_TRACE=gpurun_out/r2_aq_trace_sink.json
run nosink STCAT_GRAD_SINK=0
gzip -f gpurun_out/r2_aq_trace_sink.json
