#!/bin/bash
# warp epilogue default + bias prefetch: parity, timeline, table, step
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x > gpurun_out/r2_x_pytest.log 2>&1; echo "pytest (default) rc=$?"; grep -v Warning gpurun_out/r2_x_pytest.log | tail -5
timeout 120 python scripts/gemm_timeline.py 2>&1 | tail -22 | tee gpurun_out/r2_x_timeline.txt
timeout 300 python scripts/bench_gemm.py --iters 30 2>&1 | tee gpurun_out/r2_x_gemm_table.txt | head -32
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_x_bench.json 2> gpurun_out/r2_x_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_x_bench.json"))
    print("step: ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roofline", d["roofline"]["frac"], d["gemm_family"]["frac_of_bf16_peak"])
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_x_bench.err").read()[-1500:])
PY
