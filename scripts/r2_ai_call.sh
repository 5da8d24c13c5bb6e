#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
timeout 300 python scripts/bench_gemm.py --iters 30 2>&1 | tee gpurun_out/r2_ai_gemm_table.txt | grep "ffn\|qk_fwd\|family"
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_ai_bench_$i.json 2> gpurun_out/r2_ai_bench_$i.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_ai_bench_$i.json"))
print("dropout 0 step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), d["clocks"])
PY
done
timeout 600 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_kernels.py tests/test_gpu_hotpath.py tests/test_gpu_train_step.py -q -m gpu -k "dropout or train" 2>&1 | tail -3
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dropout 0.1 > gpurun_out/r2_ai_bench_dropout01.json 2> gpurun_out/r2_ai_bench_dropout01.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_ai_bench_dropout01.json"))
print("dropout 0.1 step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), d["clocks"])
PY
