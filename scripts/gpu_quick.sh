#!/bin/bash
# quick loop: the kernel test files given in $TESTS, then a profiled bf16 bench
set -u
mkdir -p gpurun_out
TESTS=${TESTS:-"tests/test_gpu_attention_tc.py"}
timeout 900 python -m pytest $TESTS -x -q > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/pytest_quick.log
tail -25 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline --profile gpurun_out/profile_bf16.md > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_bf16.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'],d['roofline']['us_per_launch'], 'graph', d['config']['cuda_graph'])
except Exception as e:
    print('no bench json',e)
PY
grep -v Warning gpurun_out/bench_bf16.err | tail -5
head -40 gpurun_out/profile_bf16.md
