#!/bin/bash
# PDL check: kernel tests eagerly, then bench with and without PDL
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warn gpurun_out/pytest_gpu.log | tail -3
for v in "STCAT_NO_PDL=1" "A=1"; do
  echo "=== $v"
  env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_pdl.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'value', d['value'], 'graph', d['config']['cuda_graph'], 'roof us', d['roofline']['us_per_launch'], 'encattn', d.get('encoder_attention'))"
  grep -E "capture unavailable|Error" gpurun_out/bench_pdl.err | head -3
done
