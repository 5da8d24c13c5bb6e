#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu > gpurun_out/r2_af_pytest.log 2>&1; echo "pytest gemm rc=$?"; grep -v Warning gpurun_out/r2_af_pytest.log | tail -12
timeout 300 python scripts/bench_gemm.py --iters 30 2>&1 | tee gpurun_out/r2_af_gemm_table.txt | grep "dec_\|family"
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_hotpath.py tests/test_gpu_train_step.py -q -m gpu > gpurun_out/r2_af_pytest2.log 2>&1; echo "pytest2 rc=$?"; grep -v Warning gpurun_out/r2_af_pytest2.log | tail -6
for sm in 1 0; do
STCAT_GEMM_SMALL=$sm timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_af_bench_small$sm.json 2> gpurun_out/r2_af_bench_small$sm.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_af_bench_small$sm.json"))
    print("small=$sm step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_af_bench_small$sm.err").read()[-1500:])
PY
done
