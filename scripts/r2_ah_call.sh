#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_attention_tc.py tests/test_gpu_train_step.py tests/test_gpu_hotpath.py tests/test_gpu_gemm_tc.py -q -m gpu -k "dropout or train or layernorm or fwd_bf16 or bwd_bf16 or relu_mask" > gpurun_out/r2_ah_pytest.log 2>&1; echo "pytest rc=$?"; grep -v Warning gpurun_out/r2_ah_pytest.log | tail -8
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dropout 0.1 --profile gpurun_out/r2_ah_profile_dropout.md > gpurun_out/r2_ah_bench_dropout01.json 2> gpurun_out/r2_ah_bench_dropout01.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_ah_bench_dropout01.json"))
    print("dropout 0.1 step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_ah_bench_dropout01.err").read()[-1500:])
PY
grep "dropout_kernel\|attn_tc\|attn_mma\|layernorm\|cast_bf16\|colsum" gpurun_out/r2_ah_profile_dropout.md | cut -c1-150
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_ah_bench.json 2> gpurun_out/r2_ah_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_ah_bench.json"))
print("dropout 0 step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1))
PY
