#!/bin/bash
# bench (bf16, graph) + per-kernel profile + trace
set -u
mkdir -p gpurun_out
if [ -n "${TESTS:-}" ]; then timeout 900 python -m pytest $TESTS -x -q > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sel.log; fi
STCAT_TRACE=gpurun_out/trace.json timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline --profile gpurun_out/profile_bf16.md > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_bf16.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'],d['roofline']['us_per_launch'], 'graph', d['config']['cuda_graph'], 'loss', d['e2e']['loss'])
except Exception as e:
    print('no bench json',e)
PY
grep -v Warning gpurun_out/bench_bf16.err | tail -5
gzip -f gpurun_out/trace.json
head -30 gpurun_out/profile_bf16.md | cut -c1-150
